"""The C-ABI library loads on a CPU-only box and exports every symbol include/lsq_b200.h
declares (no compute calls here - those are the -m gpu tests)."""
import ctypes
import re
import subprocess

from conftest import PKG, ROOT


def _declared():
    text = (ROOT / "include" / "lsq_b200.h").read_text()
    return sorted(set(re.findall(r"LSQB200_API\s+[\w\s\*]+?\b(lsqb200_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    names = _declared()
    for must in ("lsqb200_fwd_tensor", "lsqb200_bwd_tensor", "lsqb200_fwd_channel", "lsqb200_bwd_channel",
                 "lsqb200_weight_init_stats", "lsqb200_plan_create", "lsqb200_plan_forward", "lsqb200_plan_backward",
                 "lsqb200_cuda_version", "lsqb200_workspace_bytes", "lsqb200_last_error"):
        assert must in names
    assert len(names) >= 17


def test_library_exports_every_declared_symbol(native_lib):
    so = PKG / "torchlsq" / "libtorchlsq_b200.so"
    raw = ctypes.CDLL(str(so))
    for name in _declared():
        assert hasattr(raw, name), f"{name} declared in include/lsq_b200.h but not exported"
    out = subprocess.check_output(["nm", "-D", "--defined-only", str(so)], text=True)
    exported = {ln.split()[-1] for ln in out.splitlines() if ln.strip()}
    leaked = [s for s in exported if not s.startswith("lsqb200_") and not s.startswith("_")]
    assert not leaked, f"unexpected exported symbols: {leaked[:5]}"


def test_ctypes_prototypes_cover_the_header(native_lib):
    from torchlsq import _cabi
    assert sorted(_cabi._PROTOTYPES) == _declared()


def test_info_calls_work_without_a_gpu(native_lib):
    from torchlsq import _cabi
    assert native_lib.lsqb200_abi_version() == _cabi.ABI_VERSION == 3
    v = native_lib.lsqb200_cuda_version()
    assert v // 1000 == 12            # CUDA 12.x toolchain (CUDA_VERSION, torchlsq.cpp:25-31 semantics)
    assert native_lib.lsqb200_workspace_bytes() >= 4096 * 4 + 16384 * 16


def test_launch_geometry_is_pure_host_logic(native_lib):
    from torchlsq import _cabi
    info = _cabi.LaunchInfo()
    # per-tensor bf16, 205 M elements: 256-bit units, many splits, CTA-sized groups
    assert native_lib.lsqb200_query_launch(1, 1, 256 * 64 * 112 * 112, 2, 1, 1, ctypes.byref(info)) == 0
    assert info.vec == 16 and info.regime == 0 and info.splits > 100 and info.grid == info.splits
    # conv weight rows (C = 512 rows of 4608 fp32): one tile per row
    assert native_lib.lsqb200_query_launch(1, 512, 4608, 0, 1, 1, ctypes.byref(info)) == 0
    assert info.vec == 8 and info.splits == 1
    # NCHW axis 1 with 7x7 maps in bf16: rows of 98 bytes cannot be vectorised
    assert native_lib.lsqb200_query_launch(64, 2048, 49, 2, 1, 1, ctypes.byref(info)) == 0
    assert info.vec == 1 and info.regime == 1
    # 14x14 bf16 rows (392 B) take 64-bit units
    assert native_lib.lsqb200_query_launch(64, 1024, 196, 2, 1, 1, ctypes.byref(info)) == 0
    assert info.vec == 4
    # misaligned base pointers: scalar path
    assert native_lib.lsqb200_query_launch(1, 1, 100000, 0, 0, 0, ctypes.byref(info)) == 0
    assert info.vec == 1
    assert native_lib.lsqb200_query_launch(-1, 1, 1, 0, 0, 1, ctypes.byref(info)) == -1
    assert native_lib.lsqb200_set_tuning(b"bwd_tile_kb=4096,max_unit_bytes=16,bwd_min_tiles_per_sm=1") == 0
    assert native_lib.lsqb200_query_launch(1, 1, 1 << 28, 0, 1, 1, ctypes.byref(info)) == 0
    assert info.vec == 4 and info.splits == (1 << 30) // (4096 * 1024)
    assert native_lib.lsqb200_set_tuning(b"bogus=1") == -1
    assert native_lib.lsqb200_set_tuning(None) == 0


def test_optimizer_entries_validate_arguments_before_touching_a_device(native_lib):
    """Argument errors of the two flat-optimizer entries are host logic: they must come back as -1 with a message on a CPU-only
    box (no launch is attempted), and n == 0 is a no-op."""
    from torchlsq import _cabi
    a = _cabi.OptimArgs(0.1, 0.0, 1.0, 0.9, 0.0, 0.9, 0.999, 1e-8, 1, 0, 0)
    buf = (ctypes.c_float * 8)()
    cnt = (ctypes.c_int32 * 8)()
    p, c = ctypes.addressof(buf), ctypes.addressof(cnt)
    for call in (lambda *r: native_lib.lsqb200_flat_optimizer_step(*r), lambda pp, gg, s1, s2, n, aa, st: native_lib.lsqb200_flat_optimizer_step_sites(pp, gg, s1, s2, c, None, n, aa, st)):
        a.kind, a.momentum = 0, 0.9
        assert call(p, p, None, None, 8, a, None) == -1 and b"momentum" in native_lib.lsqb200_last_error()
        a.kind = 7
        assert call(p, p, None, None, 8, a, None) == -1 and b"kind" in native_lib.lsqb200_last_error()
        a.kind = 1
        assert call(p, p, p, None, 8, a, None) == -1 and b"Adam" in native_lib.lsqb200_last_error()
        a.beta1 = 1.5
        assert call(p, p, p, p, 8, a, None) == -1 and b"betas" in native_lib.lsqb200_last_error()
        a.beta1 = 0.9
        assert call(p, p, p, p, -1, a, None) == -1
        assert call(None, None, None, None, 0, a, None) == 0
        assert call(None, p, p, p, 8, a, None) == -1 and b"NULL" in native_lib.lsqb200_last_error()
    a.kind = 1
    assert native_lib.lsqb200_flat_optimizer_step_sites(p, p, p, p, None, None, 8, a, None) == -1      # the step counts are not optional


def test_default_tiles_are_the_in_step_sizes(native_lib):
    """The per-tensor tile sizes were chosen by an in-step sweep (profiles/r2_tile_sweep.md): 32 KB forward and 256 KB backward of
    one operand, the tile count rounded up to whole waves of resident CTAs and the tile then to whole CTA iterations.  A 205 M-element
    bf16 site: 12 544 forward tiles of exactly 32 KB, 1731 backward tiles of 232 KB (3 waves of 592)."""
    from torchlsq import _cabi
    assert native_lib.lsqb200_set_tuning(None) == 0
    info = _cabi.LaunchInfo()
    n = 256 * 64 * 112 * 112
    assert native_lib.lsqb200_query_launch(1, 1, n, 2, 0, 1, ctypes.byref(info)) == 0
    assert info.splits == 12544 and info.grid == info.splits
    assert native_lib.lsqb200_query_launch(1, 1, n, 2, 1, 1, ctypes.byref(info)) == 0
    assert info.splits == 1731 and 2 * 592 < info.splits <= 3 * 592
