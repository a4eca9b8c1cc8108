"""CPU tests of the host side: the C-ABI library loads and exports what include/nnic.h declares, the
constants it bakes in match NumPy, the weight containers keep the Keras layouts, and nothing in the
product imports the oracle or falls back to the CPU."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neural_network_image_compression_b200")
HEADER = os.path.join(ROOT, "include", "nnic.h")
LIB = os.path.join(PKG, "libnnic.so")


def header_functions():
    src = open(os.path.join(ROOT, "include", "nnic.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nnic_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(nn):
    from neural_network_image_compression_b200 import _lib
    lib = nn.load_library()
    declared = header_functions()
    assert len(declared) >= 18
    bound = {name for name, _r, _a in _lib.SYMBOLS}
    assert set(declared) == bound, (set(declared) ^ bound)
    exported = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\sT\s{name}$", exported, flags=re.M), f"{name} not exported"
        assert isinstance(getattr(lib, name), ctypes._CFuncPtr)
    assert lib.nnic_version().startswith(b"nnic-b200")


def test_library_is_sm100a_tensor_core_code():
    """The shipped binary holds sm_100a SASS with tcgen05 MMA / TMEM load / TMA instructions."""
    from neural_network_image_compression_b200 import _lib
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in out, mnemonic


def test_colour_constants_match_reference_definition(nn):
    """utils.py:7-9: ycbcr_kernel, np.linalg.inv(ycbcr_kernel), ycbcr_off -- rounded to fp32."""
    k, kinv, off = nn.colour_constants()
    K = np.array([[0.299, 0.587, 0.114], [-0.16874, -0.33126, 0.5], [0.5, -0.41869, -0.08131]])
    assert np.array_equal(k, K.astype(np.float32))
    assert np.array_equal(kinv, np.linalg.inv(K).astype(np.float32))
    assert np.array_equal(off, np.array([0, 0.5, 0.5], np.float32))


def test_weight_layouts_and_counts(nn, tmp_path):
    Wt = nn.weights
    enc = Wt.glorot_uniform("encoder", 1)
    dec = Wt.glorot_uniform("decoder", 2)
    assert enc["conv2/kernel"].shape == (5, 5, 32, 64)          # Conv2D [kh,kw,Cin,Cout]
    assert dec["dconv1/kernel"].shape == (5, 5, 64, 32)         # Conv2DTranspose [kh,kw,Cout,Cin]
    assert dec["dconv8/kernel"].shape == (5, 5, 1, 64)
    n_enc = sum(v.size for v in enc.values())
    n_dec = sum(v.size for v in dec.values())
    assert (n_enc, n_dec) == (177184, 229185)
    lim = np.sqrt(6.0 / ((32 + 64) * 25))
    assert np.abs(enc["conv2/kernel"]).max() <= lim and np.all(enc["conv2/bias"] == 0)
    Wt.check_weight_set("encoder", enc)
    bad = dict(enc); bad["conv3/kernel"] = np.zeros((3, 3, 64, 32), np.float32)
    with pytest.raises(ValueError):
        Wt.check_weight_set("encoder", bad)
    path = str(tmp_path / "encY.npz")
    Wt.save_npz(path, enc)
    back = Wt.load_npz(path)
    assert all(np.array_equal(back[k], enc[k]) for k in enc)


def test_no_cpu_fallback_without_gpu(nn):
    """Without a GPU the product raises; it never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nn.NnicError):
        nn.Encoder(0)
    with pytest.raises(nn.NnicError):
        nn.Handle(0)


def test_product_does_not_import_oracle():
    for dirpath, _d, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text, f"{f} mentions the oracle"


def test_shard_ranges_partition(nn):
    sr = nn.dist.shard_range
    for n in (0, 1, 7, 8, 65536, 65537):
        for g in (1, 2, 3, 8):
            rs = [sr(n, r, g) for r in range(g)]
            assert rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(g - 1))
            sizes = [hi - lo for lo, hi in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sr(4, 2, 2)


def test_div255_refinement():
    """tc_conv1.cu computes x.astype(float32)/255 (encoder.py:39) as q = RN(x*RN(1/255)) plus one residual step;
    the result must equal the IEEE quotient for every byte value (emulated here with exact double products)."""
    x = np.arange(256, dtype=np.float32)
    ref = (x / np.float32(255)).astype(np.float32)
    y = np.float32(1) / np.float32(255)
    assert float(y) == float.fromhex("0x1.010102p-8")
    q = (x * y).astype(np.float32)
    r = x.astype(np.float64) - q.astype(np.float64) * 255.0          # fma(-q, 255, x): exact, fits in fp32
    assert np.array_equal(r, r.astype(np.float32).astype(np.float64))
    q2 = (q.astype(np.float64) + r * np.float64(y)).astype(np.float32)
    assert np.array_equal(q2, ref)
    assert not np.array_equal(q, ref)      # the plain reciprocal multiply is not enough


# ---- file side of the codec: utils.py:30-62,85-120, training.py:12-21 -------------------------
def test_pack_unpack_latent_against_oracle(nn):
    from oracle import nnic_oracle as O
    rng = np.random.default_rng(5)
    for n, h, w in ((1, 1, 1), (2, 3, 5), (3, 8, 12)):
        lat = rng.integers(0, 256, size=(n, h, w, 96), dtype=np.uint8)
        pic = nn.pack_latent(lat)
        assert pic.shape == (n, 4 * h, 8 * w, 3) and pic.dtype == np.uint8
        assert np.array_equal(pic, O.pack_latent(lat))
        assert np.array_equal(nn.unpack_latent(pic), lat)
        assert np.array_equal(O.unpack_latent(pic), lat)
    with pytest.raises(ValueError):
        nn.pack_latent(np.zeros((1, 2, 2, 32), np.uint8))
    with pytest.raises(ValueError):
        nn.unpack_latent(np.zeros((1, 6, 8, 3), np.uint8))


def test_read_dataset_and_save_img(nn, tmp_path):
    from PIL import Image
    rng = np.random.default_rng(6)
    d = tmp_path / "set"
    d.mkdir()
    imgs = {name: rng.integers(0, 256, size=(16, 24, 3), dtype=np.uint8) for name in ("b", "a", "c.x")}
    for name, im in imgs.items():
        nn.save_img(im, str(d), name)
    Image.fromarray(rng.integers(0, 256, size=(16, 24), dtype=np.uint8)).save(d / "grey.png")   # skipped: not colour
    (d / "notes.txt").write_text("not an image")
    x, names = nn.read_dataset(str(d))
    assert names == ["a", "b", "c.x"]                      # sorted file names, extension stripped
    assert x.shape == (3, 16, 24, 3) and x.dtype == np.uint8
    for i, name in enumerate(names):
        assert np.array_equal(x[i], imgs[name])
    nn.save_img(rng.integers(0, 256, size=(8, 8, 3), dtype=np.uint8), str(d), "d")              # another size: ragged set
    x, names = nn.read_dataset(str(d))
    assert isinstance(x, list) and [a.shape for a in x] == [(1, 16, 24, 3)] * 3 + [(1, 8, 8, 3)]
    with pytest.raises(AssertionError):
        nn.save_img(np.full((4, 4, 3), 0.5), str(d), "frac")


def test_get_bpp_follows_the_reference_definition(nn):
    """training.py:14-21: 8 * PNG bytes of the [4h, 8w] byte picture of each plane / pixels of that picture."""
    rng = np.random.default_rng(7)
    flat = np.zeros((2, 8, 12, 32), np.float32)
    noise = rng.integers(0, 256, size=(2, 8, 12, 32)).astype(np.float32)
    b0, b1 = nn.get_bpp(flat), nn.get_bpp(noise)
    assert b0.shape == (2, 1) and b0.dtype == np.float32
    assert (b1 > 7.0).all() and (b0 < 1.0).all()           # incompressible noise costs ~8 bits per byte (+ header)
    pic = np.round(noise[0]).astype(np.uint8).reshape(32, 96)
    assert b1[0, 0] == np.float32(8.0 * nn.container.png_size(pic) / (32 * 96))
    assert np.allclose(nn.get_bpp(noise, tot_pixels_compressed=64 * 96), b1 * (32 * 96) / (64 * 96))


def test_header_is_plain_c_and_links(tmp_path):
    """include/nnic.h must be usable from C (the drop-in boundary is a C ABI): compile a C99 translation unit that
    references every entry point and link it against libnnic.so (no compute call is made)."""
    names = re.findall(r"^\s*(?:int|void|const char\*|uint64_t|size_t|long long)\s+(nnic_\w+)\s*\(", open(HEADER).read(), re.M)
    assert len(names) >= 20
    src = tmp_path / "use_nnic.c"
    src.write_text('#include "nnic.h"\n#include <stdio.h>\ntypedef void (*fn_t)(void);\nint main(void) {\n  fn_t fns[] = {'
                   + ", ".join(f"(fn_t){n}" for n in names)
                   + '};\n  printf("%s %d\\n", nnic_version(), (int)(sizeof fns / sizeof fns[0]));\n  return 0;\n}\n')
    exe = tmp_path / "use_nnic"
    lib_dir = os.path.dirname(LIB)
    cmd = ["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.dirname(HEADER), str(src), "-o", str(exe),
           "-L", lib_dir, "-lnnic", f"-Wl,-rpath,{lib_dir}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("nnic-b200"), r.stdout + r.stderr


def test_tensorflow_checkpoint_reader(nn, tmp_path):
    """tfbundle.py reads `<prefix>.index` / `.data-00000-of-00001` (Keras save_weights, utils.py:26-28) without
    TensorFlow.  Fixtures come from tests/tf_bundle_writer.py (same published layout, many small data blocks, prefix
    compressed keys, CRCs); both variable-naming styles; corruption is detected."""
    from tf_bundle_writer import write_bundle
    from neural_network_image_compression_b200 import tfbundle, weights as Wt
    for kind, style in (("encoder", "attr"), ("decoder", "indexed")):
        w = Wt.glorot_uniform(kind, 3, 1.3, 0.05)
        names = [layer[0] for layer in Wt.layers_of(kind)]
        tensors = {}
        for i, name in enumerate(names):
            base = name if style == "attr" else f"layer_with_weights-{i}"
            for var in ("kernel", "bias"):
                tensors[f"{base}/{var}/.ATTRIBUTES/VARIABLE_VALUE"] = w[f"{name}/{var}"]
        tensors["optimizer/iter/.ATTRIBUTES/VARIABLE_VALUE"] = np.array(7, np.int64)       # unrelated entries are ignored
        prefix = str(tmp_path / f"{kind}Y")
        write_bundle(prefix, tensors, block_bytes=120)
        idx = tfbundle.read_index(prefix + ".index")
        assert list(idx) == sorted(idx) and "" in idx and "_CHECKPOINTABLE_OBJECT_GRAPH" in idx
        got = tfbundle.keras_weights(prefix, names)
        assert set(got) == set(w)
        for k in w:
            assert got[k].dtype == np.float32 and np.array_equal(got[k], w[k])
        Wt.check_weight_set(kind, got)
    # a flipped byte in the data file is caught by the tensor checksum, one in the index by the block checksum
    data = prefix + ".data-00000-of-00001"
    raw = bytearray(open(data, "rb").read()); raw[100] ^= 0x40
    open(data, "wb").write(raw)
    with pytest.raises(ValueError, match="checksum"):
        tfbundle.read_bundle(prefix)
    raw = bytearray(open(prefix + ".index", "rb").read()); raw[20] ^= 0x01
    open(prefix + ".index", "wb").write(raw)
    with pytest.raises(ValueError):
        tfbundle.read_index(prefix + ".index")
    with pytest.raises(ValueError, match="magic"):
        open(str(tmp_path / "bad.index"), "wb").write(b"\\x00" * 64)
        tfbundle.read_index(str(tmp_path / "bad.index"))
    assert tfbundle.mask_crc(tfbundle.crc32c(b"123456789")) == (((0xE3069283 >> 15) | (0xE3069283 << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def test_tensorflow_checkpoint_reader_snappy_and_shards(tmp_path):
    """The table blocks of a bundle index may be Snappy-compressed (LevelDB block type 1) and the tensors may be spread over
    several `.data-NNNNN-of-MMMMM` files: both are read; the Snappy decoder is also checked on its own element kinds
    (long literals, 1- and 2-byte-offset copies, overlapping copies) and on malformed input."""
    from tf_bundle_writer import snappy_compress, write_bundle
    from neural_network_image_compression_b200 import tfbundle, weights as Wt
    rng = np.random.default_rng(5)
    for blob in (b"", b"a", b"abcd" * 5000, bytes(rng.integers(0, 4, size=70000, dtype=np.uint8)), bytes(rng.integers(0, 256, size=300, dtype=np.uint8)),
                 b"x" * 100000, (b"conv1/kernel/.ATTRIBUTES/VARIABLE_VALUE" + bytes(range(40))) * 60):
        packed = snappy_compress(blob)
        assert tfbundle.snappy_uncompress(packed) == blob
        if len(blob) > 1000:
            assert len(packed) < len(blob)
    assert tfbundle.snappy_uncompress(bytes([5, 0 << 2, ord("a"), 2 | (3 << 2), 1, 0])) == b"aaaaa"     # overlapping copy = run
    for bad in (bytes([4, 2 | (3 << 2), 9, 0]), bytes([9, 0 << 2, ord("a")]), bytes([3, 8 << 2, ord("a")])):
        with pytest.raises(ValueError, match="snappy"):
            tfbundle.snappy_uncompress(bad)
    w = Wt.glorot_uniform("decoder", 9, 1.1, 0.05)
    names = [layer[0] for layer in Wt.layers_of("decoder")]
    tensors = {f"{n}/{v}/.ATTRIBUTES/VARIABLE_VALUE": w[f"{n}/{v}"] for n in names for v in ("kernel", "bias")}
    for shards, snappy in ((1, True), (3, False), (2, True)):
        prefix = str(tmp_path / f"dec_{shards}_{int(snappy)}Y")
        write_bundle(prefix, tensors, block_bytes=150, num_shards=shards, snappy=snappy)
        assert os.path.exists(f"{prefix}.data-{shards - 1:05d}-of-{shards:05d}")
        got = tfbundle.keras_weights(prefix, names)
        for k in w:
            assert np.array_equal(got[k], w[k]), (shards, snappy, k)
    raw = bytearray(open(prefix + ".index", "rb").read()); raw[30] ^= 0x10
    open(prefix + ".index", "wb").write(raw)
    with pytest.raises(ValueError):
        tfbundle.read_index(prefix + ".index")


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` needs no GPU: exactly one JSON line on stdout with the contract's keys."""
    import json
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c3",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "MP/s" and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_c_example_compiles_as_pedantic_c99(tmp_path):
    """examples/c_abi_demo.c (run on the GPU by tests/test_gpu_parity.py) builds against the header and the library."""
    lib_dir = os.path.dirname(LIB)
    r = subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.dirname(HEADER),
                        os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L", lib_dir, "-lnnic", f"-Wl,-rpath,{lib_dir}", "-lm",
                        "-o", str(tmp_path / "demo")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_c_glorot_matches_numpy():
    """nnic_glorot_uniform / nnic_init_random (include/nnic.h) restate NumPy's default_rng(seed) stream in C++: the
    networks a C caller gets are bit-identical to weights.glorot_uniform(kind, seed), i.e. to the oracle's weight sets.
    Host code only: runs without a GPU."""
    import ctypes as C
    import neural_network_image_compression_b200 as nn
    from neural_network_image_compression_b200 import weights as Wt
    lib = nn.load_library()
    for set_index, kind, seed, gain, br in ((0, "encoder", 11, 1.0, 0.0), (1, "encoder", 12, 1.6, 0.05),
                                            (2, "decoder", 13, 1.0, 0.0), (3, "decoder", 14, 1.6, 0.05),
                                            (0, "encoder", 0, 1.0, 0.0), (3, "decoder", (1 << 40) + 5, 0.5, 1.0)):
        want = Wt.glorot_uniform(kind, seed, gain, br)
        got = Wt.glorot_uniform_native(kind, seed, gain, br, set_index)
        assert got.keys() == want.keys()
        for k in want:
            assert got[k].dtype == np.float32 and got[k].shape == want[k].shape and np.array_equal(got[k], want[k]), (set_index, seed, k)
    k, cin, cout = C.c_int(), C.c_int(), C.c_int()
    for set_index, kind in ((0, "encoder"), (2, "decoder")):
        for li, (_name, kk, _s, ci, co) in enumerate(Wt.layers_of(kind)):
            assert lib.nnic_layer_shape(set_index, li, C.byref(k), C.byref(cin), C.byref(cout)) == 0
            assert (k.value, cin.value, cout.value) == (kk, ci, co)
    assert lib.nnic_layer_shape(4, 0, None, None, None) == -1 and lib.nnic_glorot_uniform(0, 1, 1.0, 0.0, None, None) == -1
