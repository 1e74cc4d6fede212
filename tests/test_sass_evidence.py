"""CPU check of the built CUDA objects: the sm_100a SASS contains the instructions DESIGN.md claims
(B200_PROFILING.md: `UBLKCP` / `SYNCS.*` prove cp.async.bulk + mbarrier, `FFMA2` the packed fp32x2
arithmetic, `REDG` the warp-reduced gradient atomics, sys-scope `LDG/REDG/STG ... .SYS` the multimem
(NVLS) operations) and nothing was compiled for another architecture.  Skipped when cuobjdump is absent."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJ = os.path.join(ROOT, "diff-gaussian-rasterization_b200", "build")
LIB = os.path.join(ROOT, "diff-gaussian-rasterization_b200", "lib", "libgsr_b200.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not os.path.exists(CUOBJDUMP), reason="cuobjdump not available")


def _sass(obj):
    return subprocess.run([CUOBJDUMP, "-sass", os.path.join(OBJ, obj)], capture_output=True, text=True).stdout


def _count(text, pattern):
    return len(re.findall(pattern, text))


def test_library_is_sm_100a_only(built):
    out = subprocess.run([CUOBJDUMP, "-lelf", LIB], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_per_gaussian_kernels_use_bulk_async_copies(built):
    fwd, bwd = _sass("preprocess_fwd.o"), _sass("preprocess_bwd.o")
    assert _count(fwd, r"UBLKCP\.S\.G") >= 1                      # global -> shared SH rows
    assert _count(bwd, r"UBLKCP\.S\.G") >= 1 and _count(bwd, r"UBLKCP\.G\.S") >= 1   # and dL/dSH rows back
    for text in (fwd, bwd):
        assert _count(text, r"SYNCS\.ARRIVE\.TRANS64") >= 1      # mbarrier arrive / expect_tx
        assert _count(text, r"SYNCS\.PHASECHK\.TRANS64\.TRYWAIT") >= 1


def test_backward_blend_uses_packed_fp32x2_and_warp_reduced_atomics(built):
    bwd = _sass("render_bwd.o")
    assert _count(bwd, r"\bFFMA2\b") >= 20 and _count(bwd, r"\bFMUL2\b") >= 20 and _count(bwd, r"\bFADD2\b") >= 5
    assert _count(bwd, r"REDG\.E\.ADD\.F32") >= 1                 # one coalesced red per (warp, Gaussian, slot)
    assert _count(bwd, r"\bATOMG\b") == 0                         # no returning atomics in the blend


def test_forward_blend_register_budget_and_packed_arithmetic(built):
    """Scalar forward kernels: 48 registers, no spills (5 CTAs of 256 threads per SM).  Packed
    quarter-list forward kernels: at most 64 registers (8 CTAs of 128 threads per SM) with at most a
    16-byte spill frame, and their blend really is FFMA2 / FMUL2."""
    out = subprocess.run([CUOBJDUMP, "-res-usage", os.path.join(OBJ, "render_fwd.o")], capture_output=True,
                         text=True).stdout
    rows = re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out)
    assert rows
    for name, reg, stack in rows:
        reg, stack = int(reg), int(stack)
        if "render_fwdq_kernel" in name:
            assert reg <= 64 and stack <= 16, (name, reg, stack)
        elif "render_fwd_kernel" in name:
            assert stack == 0 and reg <= 64, (name, reg, stack)
    assert any("render_fwdq_kernel" in n for n, _, _ in rows)
    fwd = _sass("render_fwd.o")
    assert _count(fwd, r"\bFFMA2\b") >= 20 and _count(fwd, r"\bFMUL2\b") >= 20


def test_nvls_kernels_use_sys_scope_multimem_operations(built):
    sh = _sass("sh_grad_views.o")
    # multimem.ld_reduce -> LDGMC (load with in-switch reduction), multimem.st -> sys-scope STG.128
    assert _count(sh, r"LDGMC\.E\.ADD\.F32x4") >= 1, "no multimem.ld_reduce in sh_grad_views.o"
    assert _count(sh, r"STG\.E\.128\.STRONG\.SYS") >= 1


def test_p2p_exchange_kernels_use_sys_scope_128_bit_accesses(built):
    """The two-shot slice all-reduce and the gather read / write the peers' replicas with 128-bit sys-scope
    loads and stores (ld / st.relaxed.sys.v4)."""
    sh = _sass("sh_grad_views.o")
    assert _count(sh, r"LDG\.E\.128\.STRONG\.SYS") >= 8 and _count(sh, r"STG\.E\.128\.STRONG\.SYS") >= 8


def test_frame_kernels_are_chained_by_programmatic_dependent_launch(built):
    """griddepcontrol.wait -> ACQBULK at the top of the dependent kernels (tile scan, per-tile sort, forward blend,
    per-Gaussian backward), griddepcontrol.launch_dependents -> PREEXIT in the kernels in front of them."""
    waits = {o: _count(_sass(o), r"\bACQBULK\b") for o in ("binning.o", "render_fwd.o", "preprocess_bwd.o")}
    triggers = {o: _count(_sass(o), r"\bPREEXIT\b") for o in ("preprocess_fwd.o", "binning.o", "render_bwd.o")}
    assert all(v >= 1 for v in waits.values()), waits
    assert all(v >= 1 for v in triggers.values()), triggers


def test_backward_blend_and_sort_register_budgets(built):
    """The default (quarter-list) backward blend kernels fit 7 CTAs of 128 threads per SM (<= 72 registers, no
    spills); the per-tile sort fits 6 CTAs of 256 threads (<= 40 registers, no spills)."""
    out = subprocess.run([CUOBJDUMP, "-res-usage", os.path.join(OBJ, "render_bwd.o")], capture_output=True, text=True).stdout
    rows = re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out)
    q7 = [(n, int(r), int(st)) for n, r, st in rows if re.search(r"render_bwdq_kernelILi[01]ELb[01]ELi7ELb0E", n)]   # (not the exact_median build)
    assert len(q7) >= 3
    for n, r, st in q7:
        assert r <= 72 and st == 0, (n, r, st)
    out = subprocess.run([CUOBJDUMP, "-res-usage", os.path.join(OBJ, "binning.o")], capture_output=True, text=True).stdout
    rows = re.findall(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out)
    srt = [(n, int(r), int(st)) for n, r, st in rows if "sort_tiles_kernel" in n]
    assert srt and all(r <= 40 and st == 0 for _, r, st in srt), srt
