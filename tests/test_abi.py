"""The C-ABI library loads and exports every symbol include/softgrip.h declares (no compute without a GPU)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

from conftest import ROOT, blob_path, pkg

HEADER = os.path.join(ROOT, "include", "softgrip.h")


@pytest.fixture(scope="module")
def libpath():
    lib = pkg("_lib")
    if not os.path.exists(lib.LIB_PATH):
        sys.path.insert(0, ROOT)
        import __graft_entry__ as g
        g.build()
    return lib.LIB_PATH


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sg_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(libpath):
    L = ctypes.CDLL(libpath)
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(L, s), "libsoftgrip.so does not export %s" % s


def test_python_signatures_cover_the_header(libpath):
    lib = pkg("_lib")
    assert sorted(lib.SIGNATURES) == declared_symbols()


def test_exports_are_c_linkage(libpath):
    out = subprocess.run(["nm", "-D", "--defined-only", libpath], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert set(declared_symbols()) <= exported


def test_model_load_and_info_need_no_gpu(libpath):
    """Plan building (level schedule, pair lists) is host code and must work on the CPU-only box."""
    lib = pkg("_lib")
    L = lib.lib()
    for name, nv, nshell, neq, nlev in (("softbox", 118, 110, 327, 53), ("softball", 226, 218, 651, 73), ("softcylinder", 200, 192, 573, 69)):
        blob = open(blob_path(name), "rb").read()
        h = ctypes.c_void_p()
        assert L.sg_model_load(blob, len(blob), ctypes.byref(h)) == 0, L.sg_last_error()
        info = lib.SgInfo()
        assert L.sg_model_info(h, ctypes.byref(info)) == 0
        assert (info.nv, info.nshell, info.neq, info.nfinger, info.nu, info.nsensordata) == (nv, nshell, neq, 8, 2, 12)
        assert info.nlevels == nlev            # Gauss-Seidel depth derived in SURVEY section 7 ("Hard parts")
        assert 0 < info.smem_bytes32 < info.smem_bytes64 <= 227 * 1024
        L.sg_model_destroy(h)


def test_bad_blob_is_an_error_not_a_crash(libpath):
    lib = pkg("_lib")
    L = lib.lib()
    h = ctypes.c_void_p()
    assert L.sg_model_load(b"garbage" * 10, 70, ctypes.byref(h)) < 0
    assert b"model blob" in L.sg_last_error() or b"missing" in L.sg_last_error()


def test_batch_create_fails_loudly_without_gpu(libpath):
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    lib = pkg("_lib")
    L = lib.lib()
    blob = open(blob_path("softbox"), "rb").read()
    h = ctypes.c_void_p()
    assert L.sg_model_load(blob, len(blob), ctypes.byref(h)) == 0
    b = ctypes.c_void_p()
    assert L.sg_batch_create(h, 4, 0, 32, ctypes.byref(b)) < 0
    assert b"no CUDA device" in L.sg_last_error()
    batched = pkg("batched")
    with pytest.raises(lib.SoftGripError):
        batched.BatchedManEnv(blob_path("softbox"), 4)


def build_c_caller(libpath, out_dir):
    """tests/cabi/rollout_host.c: a plain C99 program against include/softgrip.h and libsoftgrip.so (no Python, no torch)."""
    exe = os.path.join(str(out_dir), "rollout_host")
    libdir = os.path.dirname(libpath)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-O1", "-I" + os.path.join(ROOT, "include"), "-o", exe,
                           os.path.join(ROOT, "tests", "cabi", "rollout_host.c"), "-L" + libdir, "-l:libsoftgrip.so",
                           "-Wl,-rpath," + libdir])
    return exe


def test_plain_c_caller_builds_and_fails_loudly_without_gpu(libpath, tmp_path):
    """The boundary is usable from C as declared (header is valid C99, symbols link), model loading needs no GPU, and a
    CPU-only box gets the explicit "no CUDA device" error instead of a fallback."""
    torch = pytest.importorskip("torch")
    exe = build_c_caller(libpath, tmp_path)
    if torch.cuda.is_available():
        pytest.skip("CUDA present: the full run is tests/test_gpu.py::test_plain_c_caller_matches_the_python_path")
    out = subprocess.run([exe, blob_path("softbox"), "4", str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=120)
    assert out.returncode == 3, (out.returncode, out.stderr)
    assert "nv 118 nshell 110 neq 327 nu 2 nsensordata 12 levels 53" in out.stdout
    assert "no CUDA device" in out.stderr and not os.path.exists(str(tmp_path / "out.bin"))


def test_integration_doc_maps_every_entry_point():
    """INTEGRATION.md's call-site map names every symbol the header declares."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [s for s in declared_symbols() if s not in doc]
    assert not missing, missing


def test_step_kernel_carries_the_tensor_memory_and_ring_instructions(libpath):
    """The in-tree library's fp32 8-lane step kernel was compiled for sm_100a with the equality rows in tensor memory
    (LDTM / STTM, allocation by UTCATOMSWS) and the contact-record ring (LDGSTS + LDGDEPBAR), and without tensor-core
    math (the path has no dense contraction to offer).  Skipped without cuobjdump."""
    import shutil
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("no cuobjdump")
    listing = subprocess.run([cuobjdump, "-lelf", libpath], capture_output=True, text=True).stdout
    assert "sm_100a" in listing, listing[:300]
    sass = subprocess.run([cuobjdump, "-sass", libpath], capture_output=True, text=True).stdout
    i = sass.find("Function : _ZN2sg15sg_step_kernel2IfLi8EEEvNS_6KArgs2IT_EE")
    assert i >= 0
    j = sass.find("Function :", i + 10)
    body = sass[i:j if j > 0 else len(sass)]
    for mnemonic in ("LDTM", "STTM", "UTCATOMSWS", "LDGSTS", "LDGDEPBAR"):
        assert mnemonic in body, mnemonic
    assert not re.search(r"\b(HMMA|UTCHMMA|UTCQMMA|UTCIMMA)\b", body)
