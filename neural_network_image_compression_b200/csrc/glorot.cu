// Keras-default initialisation of one network, for callers of the C ABI that have no checkpoint:
// glorot-uniform kernels, zero biases (what a freshly constructed Encoder() / Decoder() of the reference holds,
// tf2_0/src/encoder.py:34-36, decoder.py:35-37 -> keras.layers.Conv2D defaults).
//
// The random stream is NumPy's: np.random.default_rng(seed) = PCG64 seeded through SeedSequence(seed), doubles drawn as
// (next64 >> 11) * 2^-53, Generator.uniform(low, high) = low + (high - low) * u.  The weight sets of the parity tests
// (neural_network_image_compression_b200/weights.py::glorot_uniform, tests/conftest.py::make_weights) come from exactly
// this stream, so nnic_init_random(h, set, seed) from C and weights.glorot_uniform(kind, seed) from Python give
// bit-identical networks (tests/test_host.py::test_c_glorot_matches_numpy, no GPU needed).  Host code only.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/nnic.h"

namespace {

// ---- numpy.random.SeedSequence (bit_generator.pyx), pool of four 32-bit words ---------------------------------
constexpr uint32_t INIT_A = 0x43b0d7e5u, MULT_A = 0x931e8875u, INIT_B = 0x8b51f9ddu, MULT_B = 0x58f38dedu;
constexpr uint32_t MIX_MULT_L = 0xca01f9ddu, MIX_MULT_R = 0x4973f715u;
constexpr int XSHIFT = 16, POOL = 4;

inline uint32_t hashmix(uint32_t value, uint32_t& hash_const) {
  value ^= hash_const;
  hash_const *= MULT_A;
  value *= hash_const;
  value ^= value >> XSHIFT;
  return value;
}
inline uint32_t mix(uint32_t x, uint32_t y) {
  uint32_t r = MIX_MULT_L * x - MIX_MULT_R * y;
  r ^= r >> XSHIFT;
  return r;
}
void seed_sequence_state(uint64_t seed, uint64_t out[4]) {
  // entropy as little-endian 32-bit words (at least one word)
  std::vector<uint32_t> ent;
  ent.push_back((uint32_t)seed);
  if (seed >> 32) ent.push_back((uint32_t)(seed >> 32));
  uint32_t pool[POOL];
  uint32_t hc = INIT_A;
  for (int i = 0; i < POOL; ++i) pool[i] = hashmix(i < (int)ent.size() ? ent[i] : 0u, hc);
  for (int s = 0; s < POOL; ++s)
    for (int d = 0; d < POOL; ++d)
      if (s != d) pool[d] = mix(pool[d], hashmix(pool[s], hc));
  for (size_t s = POOL; s < ent.size(); ++s)
    for (int d = 0; d < POOL; ++d) pool[d] = mix(pool[d], hashmix(ent[s], hc));
  // generate_state(4, uint64) = 8 uint32 words, pairs viewed little-endian
  uint32_t words[8];
  uint32_t hb = INIT_B;
  for (int i = 0; i < 8; ++i) {
    uint32_t v = pool[i % POOL];
    v ^= hb;
    hb *= MULT_B;
    v *= hb;
    v ^= v >> XSHIFT;
    words[i] = v;
  }
  for (int i = 0; i < 4; ++i) out[i] = (uint64_t)words[2 * i] | ((uint64_t)words[2 * i + 1] << 32);
}

// ---- PCG64 (XSL-RR 128/64), numpy/random/src/pcg64 ---------------------------------------------------------------
struct Pcg64 {
  unsigned __int128 state, inc;
  static unsigned __int128 mult() { return ((unsigned __int128)2549297995355413924ULL << 64) | 4865540595714422341ULL; }
  void step() { state = state * mult() + inc; }
  explicit Pcg64(uint64_t seed) {
    uint64_t s[4];
    seed_sequence_state(seed, s);
    const unsigned __int128 initstate = ((unsigned __int128)s[0] << 64) | s[1];
    const unsigned __int128 initseq = ((unsigned __int128)s[2] << 64) | s[3];
    state = 0;
    inc = (initseq << 1) | 1;
    step();
    state += initstate;
    step();
  }
  uint64_t next64() {
    step();
    const uint64_t hi = (uint64_t)(state >> 64), lo = (uint64_t)state;
    const uint64_t x = hi ^ lo;
    const unsigned rot = (unsigned)(state >> 122);
    return (x >> rot) | (x << ((64 - rot) & 63));
  }
  double next_double() { return (double)(next64() >> 11) * (1.0 / 9007199254740992.0); }
};

struct Spec { int k, cin, cout; };
const Spec kEncSpec[5] = {{5, 1, 32}, {5, 32, 64}, {3, 64, 64}, {3, 64, 64}, {5, 64, 32}};   // encoder.py:10-17
const Spec kDecSpec[5] = {{5, 32, 64}, {3, 64, 64}, {3, 64, 64}, {5, 64, 64}, {5, 64, 1}};   // decoder.py:10-17

}  // namespace

extern "C" {

int nnic_layer_shape(int set, int layer, int* ksize, int* cin, int* cout) {
  if (set < 0 || set > 3 || layer < 0 || layer >= NNIC_LAYERS_PER_NET) return NNIC_ERR_INVALID_ARG;
  const Spec& s = set < 2 ? kEncSpec[layer] : kDecSpec[layer];
  if (ksize) *ksize = s.k;
  if (cin) *cin = s.cin;
  if (cout) *cout = s.cout;
  return NNIC_OK;
}

// kernels: the five layers' kernels back to back in their Keras layouts ([kh,kw,Cin,Cout] for the encoder sets,
// [kh,kw,Cout,Cin] for the decoder sets -- both are k*k*cin*cout values per layer); biases: the five [Cout] vectors back to back.
int nnic_glorot_uniform(int set, uint64_t seed, double gain, double bias_range, float* kernels, float* biases) {
  if (set < 0 || set > 3 || !kernels || !biases) return NNIC_ERR_INVALID_ARG;
  const Spec* specs = set < 2 ? kEncSpec : kDecSpec;
  Pcg64 rng(seed);
  for (int l = 0; l < NNIC_LAYERS_PER_NET; ++l) {
    const Spec& s = specs[l];
    const double limit = std::sqrt(6.0 / (double)((s.cin + s.cout) * s.k * s.k));
    const double low = -limit, range = limit - low;
    const size_t nk = (size_t)s.k * s.k * s.cin * s.cout;
    for (size_t i = 0; i < nk; ++i) {
      volatile double prod = range * rng.next_double();      // separate roundings, as NumPy computes low + range * u
      const double v = low + prod;
      volatile double scaled = v * gain;
      kernels[i] = (float)scaled;
    }
    kernels += nk;
    if (bias_range > 0.0) {
      const double blow = -bias_range, brange = bias_range - blow;
      for (int i = 0; i < s.cout; ++i) {
        volatile double prod = brange * rng.next_double();
        biases[i] = (float)(blow + prod);
      }
    } else {
      for (int i = 0; i < s.cout; ++i) biases[i] = 0.0f;
    }
    biases += s.cout;
  }
  return NNIC_OK;
}

int nnic_init_random_scaled(nnic_t* h, int set, uint64_t seed, double gain, double bias_range) {
  if (!h || set < 0 || set > 3) return NNIC_ERR_INVALID_ARG;
  const Spec* specs = set < 2 ? kEncSpec : kDecSpec;
  size_t nk = 0, nb = 0;
  for (int l = 0; l < NNIC_LAYERS_PER_NET; ++l) { nk += (size_t)specs[l].k * specs[l].k * specs[l].cin * specs[l].cout; nb += specs[l].cout; }
  std::vector<float> kern(nk), bias(nb);
  int rc = nnic_glorot_uniform(set, seed, gain, bias_range, kern.data(), bias.data());
  if (rc != NNIC_OK) return rc;
  const float* kp = kern.data();
  const float* bp = bias.data();
  for (int l = 0; l < NNIC_LAYERS_PER_NET; ++l) {
    rc = nnic_set_weights(h, set, l, kp, bp);
    if (rc != NNIC_OK) return rc;
    kp += (size_t)specs[l].k * specs[l].k * specs[l].cin * specs[l].cout;
    bp += specs[l].cout;
  }
  return NNIC_OK;
}

int nnic_init_random(nnic_t* h, int set, uint64_t seed) { return nnic_init_random_scaled(h, set, seed, 1.0, 0.0); }

}  // extern "C"
