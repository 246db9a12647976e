"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/kzgb200.h declares,
fails loudly (KZGB200_INTERNAL_ERROR, no CPU fallback) when there is no GPU, and the C++ mirror include/kzg_rs.hpp
compiles and links against it.  No compute calls are made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "kzg_rs_b200", "libkzgb200.so")


@pytest.fixture(scope="module")
def dll():
    if not os.path.exists(LIB):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "kzg_rs_b200", "csrc")])
    return C.CDLL(LIB)


def test_every_declared_symbol_is_exported(dll):
    hdr = open(os.path.join(ROOT, "include", "kzgb200.h")).read()
    names = sorted(set(re.findall(r"\b(kzgb200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 16
    for n in names:
        assert hasattr(dll, n), "libkzgb200.so does not export %s" % n


def test_no_cpu_fallback_without_gpu(dll):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = C.c_void_p()
    dll.kzgb200_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_char_p, C.c_size_t]
    assert dll.kzgb200_create(C.byref(ctx), 0, bytes(192), 192) == 2      # KZGB200_INTERNAL_ERROR
    assert not ctx.value
    assert dll.kzgb200_create(C.byref(ctx), 0, bytes(10), 10) == 5        # KZGB200_INVALID_SETUP


def test_pipeline_argument_checks(dll):
    """Streaming front-end: argument validation needs no device; creation fails loudly without one."""
    h, ok, t = C.c_void_p(), C.c_int(), C.c_uint64()
    dll.kzgb200_pipeline_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_char_p, C.c_size_t, C.c_int]
    dll.kzgb200_pipeline_wait.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(C.c_int)]
    dll.kzgb200_pipeline_depth.argtypes = [C.c_void_p]
    assert dll.kzgb200_pipeline_create(C.byref(h), 0, bytes(192), 192, 0) == 1          # depth out of range
    assert dll.kzgb200_pipeline_create(C.byref(h), 0, bytes(192), 192, 9) == 1
    assert dll.kzgb200_pipeline_wait(None, 1, C.byref(ok)) == 1
    assert dll.kzgb200_pipeline_depth(None) == 0
    import torch
    if not torch.cuda.is_available():
        assert dll.kzgb200_pipeline_create(C.byref(h), 0, bytes(192), 192, 2) == 2      # no device -> InternalError, no fallback
        assert not h.value


def test_python_mirror_types_and_errors():
    import kzg_rs_b200 as K
    assert len(K.Bytes32.from_slice(bytes(32))) == 32 and len(K.Bytes48.from_slice(bytes(48))) == 48   # dtypes.rs:61-71
    for cls, n in ((K.Bytes32, 31), (K.Bytes48, 49), (K.Blob, 131071)):
        with pytest.raises(K.KzgError) as e:
            cls.from_slice(bytes(n))
        assert e.value.kind == "InvalidBytesLength"
    with pytest.raises(K.KzgError) as e:
        K.Bytes32.from_hex("0xzz")
    assert e.value.kind == "InvalidHexFormat"
    s = K.KzgSettings.load_trusted_setup_file()
    assert len(s.g1_lagrange_bytes) == 4096 * 48 and len(s.g2_monomial_bytes) == 65 * 96


def test_product_package_does_not_touch_the_oracle():
    """The product path must never import / link the oracle."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "kzg_rs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                for needle in ("import oracle", "from oracle", "libkzg_oracle", "oracle/", "kzgo_"):
                    assert needle not in src, "%s references the oracle (%s)" % (os.path.join(dirpath, f), needle)
    out = subprocess.run(["ldd", LIB], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_cpp_mirror_compiles_and_links(tmp_path, dll):
    src = tmp_path / "use_mirror.cpp"
    src.write_text(r'''
#include "kzg_rs.hpp"
#include <cstdio>
int main(int argc, char** argv) {
    using namespace kzg_rs;
    auto b48 = Bytes48::from_slice(nullptr, 47);
    if (!b48.is_err() || b48.unwrap_err().kind != KzgError::InvalidBytesLength) return 1;
    auto settings = KzgSettings::load_trusted_setup_file(argv[1]);
    if (settings.is_err()) { std::puts("no-gpu"); return 0; }       // fails loudly without a GPU: that is the contract
    std::vector<Blob> blobs; std::vector<Bytes48> cs, ps;
    auto r = KzgProof::verify_blob_kzg_proof_batch(blobs, cs, ps, settings.unwrap());   // n = 0 -> Ok(true)
    auto pipe = BatchPipeline::create(settings.unwrap(), 2);
    if (pipe.is_err()) { std::puts("unexpected"); return 0; }
    auto t = pipe.unwrap().submit(blobs, cs, ps);
    auto w = pipe.unwrap().wait(t.unwrap());
    std::puts(r.is_ok() && r.unwrap() && w.is_ok() && w.unwrap() ? "ok-true" : "unexpected");
    return 0;
}
''')
    exe = tmp_path / "use_mirror"
    subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", os.path.dirname(LIB), "-lkzgb200", "-Wl,-rpath," + os.path.dirname(LIB)])
    out = subprocess.run([str(exe), os.path.join(ROOT, "kzg_rs_b200", "data", "mainnet_setup.bin")], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() in ("no-gpu", "ok-true")


def test_host_transcript_sha256_both_code_paths(dll):
    """The batch transcript (compute_r_powers, reference src/kzg_proof.rs:291-348) is hashed by the library's host code:
    SHA-NI when the CPU has it, portable otherwise.  Both against hashlib, over the padding edge cases and a transcript-sized message."""
    import hashlib
    import random
    dll.kzgb200_host_sha256.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_int]
    rnd = random.Random(7)
    out = C.create_string_buffer(32)
    for n in list(range(0, 130)) + [32 + 160 * 64, 32 + 160 * 1000 + 0, 1 << 20]:
        msg = rnd.randbytes(n)
        for portable in (0, 1):
            used = dll.kzgb200_host_sha256(msg, n, out, portable)
            assert out.raw == hashlib.sha256(msg).digest(), (n, portable)
            assert not (portable and used)
