/* Using libnnic.so from plain C (host buffers): encode -> rate -> decode of one synthetic image.
 *
 *   gcc -std=c99 -O2 -Iinclude examples/c_abi_demo.c -Lneural_network_image_compression_b200 -lnnic \
 *       -Wl,-rpath,$PWD/neural_network_image_compression_b200 -lm -o /tmp/c_abi_demo && /tmp/c_abi_demo
 *
 * The weights are a Keras-style glorot-uniform draw from a small LCG (any real use calls nnic_set_weights with the arrays
 * of a trained checkpoint, in the Keras layouts).  Prints bits per pixel, PSNR and FNV-1a checksums of the latent and the
 * reconstruction; tests/test_gpu_parity.py runs it and recomputes the same numbers through the Python layer.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "nnic.h"

static uint32_t lcg_state = 12345u;
static float lcg_uniform(void) {                 /* [0, 1) */
  lcg_state = lcg_state * 1664525u + 1013904223u;
  return (float)(lcg_state >> 8) * (1.0f / 16777216.0f);
}
static uint64_t fnv1a(const uint8_t* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}
#define CHECK(call) do { int rc_ = (call); if (rc_ != NNIC_OK) { \
  fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, nnic_last_error(h)); return 1; } } while (0)

int main(void) {
  /* (ksize, cin, cout) per layer; Conv2D kernels are [k,k,cin,cout], Conv2DTranspose kernels [k,k,cout,cin] */
  static const int enc[5][3] = {{5, 1, 32}, {5, 32, 64}, {3, 64, 64}, {3, 64, 64}, {5, 64, 32}};
  static const int dec[5][3] = {{5, 32, 64}, {3, 64, 64}, {3, 64, 64}, {5, 64, 64}, {5, 64, 1}};
  enum { N = 2, H = 64, W = 96 };
  nnic_t* h = NULL;
  if (nnic_create(0, &h) != NNIC_OK) { fprintf(stderr, "nnic_create: %s\n", nnic_last_error(NULL)); return 1; }
  for (int set = 0; set < 4; ++set)
    for (int l = 0; l < NNIC_LAYERS_PER_NET; ++l) {
      const int* s = set < 2 ? enc[l] : dec[l];
      const size_t nk = (size_t)s[0] * s[0] * s[1] * s[2];
      const float limit = 1.6f * sqrtf(6.0f / (float)((s[1] + s[2]) * s[0] * s[0]));
      float* k = (float*)malloc(nk * sizeof(float));
      float* b = (float*)malloc((size_t)s[2] * sizeof(float));
      for (size_t i = 0; i < nk; ++i) k[i] = (2.0f * lcg_uniform() - 1.0f) * limit;
      for (int i = 0; i < s[2]; ++i) b[i] = (2.0f * lcg_uniform() - 1.0f) * 0.05f;
      CHECK(nnic_set_weights(h, set, l, k, b));
      free(k); free(b);
    }
  static uint8_t rgb[N * H * W * 3], rec[N * H * W * 3], latent[N * (H / 8) * (W / 8) * 96];
  for (int n = 0; n < N; ++n)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        for (int c = 0; c < 3; ++c)
          rgb[((n * H + y) * W + x) * 3 + c] = (uint8_t)(128.0 + 100.0 * sin(0.11 * x + 0.07 * y * (c + 1) + n) + 20.0 * lcg_uniform());
  uint32_t hist[N * 3 * 256];
  float entropy[N * 3], bpp[N];
  uint64_t hist_global[3 * 256] = {0};
  CHECK(nnic_encode_rate(h, rgb, N, H, W, latent, hist, entropy, bpp, hist_global, NNIC_MEM_HOST, NULL));
  CHECK(nnic_decode(h, latent, N, H / 8, W / 8, rec, NULL, NNIC_MEM_HOST, NULL));
  double se = 0.0;
  for (size_t i = 0; i < sizeof rgb; ++i) { const double d = (double)rgb[i] - (double)rec[i]; se += d * d; }
  uint64_t total = 0;
  for (int i = 0; i < 3 * 256; ++i) total += hist_global[i];
  printf("%s\n", nnic_version());
  printf("bpp %.6f %.6f  entropy[0] %.6f %.6f %.6f\n", bpp[0], bpp[1], entropy[0], entropy[1], entropy[2]);
  printf("psnr %.4f dB  symbols %llu  launches %llu\n", 10.0 * log10(255.0 * 255.0 / (se / (double)sizeof rgb)),
         (unsigned long long)total, (unsigned long long)nnic_launch_count(h));
  printf("latent_fnv %016llx recon_fnv %016llx\n", (unsigned long long)fnv1a(latent, sizeof latent),
         (unsigned long long)fnv1a(rec, sizeof rec));
  /* Second pass with the Keras-default networks a fresh Encoder() / Decoder() of the reference holds: nnic_init_random draws
   * them inside the library (NumPy default_rng stream; seeds 11..14 are the parity tests' "default" weight sets), and the
   * per-feature-channel table of nnic_rate_channels must add up to the per-plane counts. */
  for (int set = 0; set < 4; ++set) CHECK(nnic_init_random(h, set, 11u + (uint64_t)set));
  uint64_t hist_global2[3 * 256] = {0};
  static uint64_t hist_ch[96 * 256];
  CHECK(nnic_encode_rate(h, rgb, N, H, W, latent, NULL, NULL, bpp, hist_global2, NNIC_MEM_HOST, NULL));
  CHECK(nnic_rate_channels(h, latent, N, H / 8, W / 8, hist_ch, NNIC_MEM_HOST, NULL));
  CHECK(nnic_decode(h, latent, N, H / 8, W / 8, rec, NULL, NNIC_MEM_HOST, NULL));
  int rows_ok = 1;
  for (int p = 0; p < 3; ++p)
    for (int b = 0; b < 256; ++b) {
      uint64_t s = 0;
      for (int c = 0; c < 32; ++c) s += hist_ch[(p * 32 + c) * 256 + b];
      if (s != hist_global2[p * 256 + b]) rows_ok = 0;
    }
  printf("default_latent_fnv %016llx default_recon_fnv %016llx channel_rows_ok %d default_bpp %.6f %.6f\n",
         (unsigned long long)fnv1a(latent, sizeof latent), (unsigned long long)fnv1a(rec, sizeof rec), rows_ok, bpp[0], bpp[1]);
  nnic_destroy(h);
  return total == sizeof latent && rows_ok ? 0 : 2;
}
