"""Static checks on the built library (cuobjdump runs without a GPU): the hot kernels are the hand-written sm_100a code
paths DESIGN.md describes -- bulk-copy (TMA) staging in the pass kernels, tcgen05 MMAs with TMEM loads in the batched GEMMs,
packed f32x2 math and MUFU.EX2 in the per-frame pass -- and keep the occupancy their launch configuration assumes."""
import re
import shutil
import subprocess

import pytest

from jaxent_b200 import _build

cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
pytestmark = pytest.mark.skipif(shutil.which(cuobjdump) is None, reason="cuobjdump not available")


@pytest.fixture(scope="module", autouse=True)
def _built():
    if _build.needs_build():
        _build.build()


@pytest.fixture(scope="module")
def resources():
    out = subprocess.run([cuobjdump, "--dump-resource-usage", _build.LIB], capture_output=True, text=True, check=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+)", out):
        res[m.group(1)] = dict(reg=int(m.group(2)), stack=int(m.group(3)), shared=int(m.group(4)))
    assert res, "no sm_100a kernels found in the library"
    return res


def _sass(pattern):
    out = subprocess.run([cuobjdump, "-sass", "-fun", pattern, _build.LIB], capture_output=True, text=True).stdout
    return re.findall(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", out, flags=re.M)


def _one(resources, *needles):
    hits = [k for k in resources if all(n in k for n in needles)]
    assert len(hits) == 1, (needles, hits)
    return hits[0]


def test_library_targets_sm_100a():
    out = subprocess.run([cuobjdump, "-lelf", _build.LIB], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in out


def test_pass_kernel_occupancy_and_bulk_copies(resources):
    k = _one(resources, "bv_pass_kernelILi1ELi0EEE")  # fp32 features, one row per thread
    assert resources[k]["reg"] <= 60  # 2 CTAs x 544 threads per SM (launch bounds of the HBM-bound configuration)
    ops = _sass(k)
    assert any(o.startswith("UBLKCP") for o in ops), "the feature tiles are staged by cp.async.bulk"
    assert any(o.startswith("SYNCS") for o in ops), "mbarrier transactions"
    assert sum(o.startswith("BAR") for o in ops) >= 4  # named barriers between the consumers and the Adam warp


def test_per_frame_pass_uses_packed_math_and_mufu(resources):
    k = _one(resources, "bv_pass_kernel_pf")
    assert resources[k]["reg"] <= 96  # 18 warps: five on two of the four sub-partitions
    ops = _sass(k)
    assert sum(o.startswith("MUFU.EX2") for o in ops) >= 72  # (T + 1) exponentials per (row, frame), two quads in flight
    assert sum(o.startswith("FFMA2") for o in ops) >= 64 and any(o.startswith("FMUL2") for o in ops)
    assert any(o.startswith("UBLKCP") for o in ops)


def test_batched_gemms_are_tcgen05(resources):
    for mode in ("ILi1E", "ILi2E"):
        k = _one(resources, "bg_gemm_kernel", mode)
        ops = _sass(k)
        assert any(o.startswith("UTCHMMA") for o in ops), "tcgen05.mma kind::f16"
        assert any(o.startswith("LDTM") for o in ops), "tcgen05.ld from tensor memory"
        assert any(o.startswith("UTCBAR") for o in ops), "tcgen05.commit -> mbarrier"
        assert any(o.startswith("UBLKCP") for o in ops), "operands staged by bulk copies"
        assert not any(o.startswith(("HMMA", "WGMMA")) for o in ops), "no mma.sync / wgmma fallback"
        assert resources[k]["stack"] == 0


def test_tail_kernels_do_not_copy_their_parameters(resources):
    # a by-reference use of the kernel parameter struct once forced a 1.7 KB local copy per thread (tail 21 -> 65 us)
    for k, r in resources.items():
        if "bv_tail_kernel" in k:
            assert r["stack"] <= 512, (k, r)


def test_per_frame_bv_gradient_sweep(resources):
    """pf_dbeta_kernel (d loss / d (bc, bh) of the per-frame mode): exponentials on the special-function unit, no spills, and
    enough registers left for two 512-thread CTAs per SM; the fold is a single small kernel"""
    for tm in ("ILi4E", "ILi8E"):
        k = _one(resources, "pf_dbeta_kernel", tm)
        assert resources[k]["stack"] == 0 and resources[k]["reg"] <= 64
        ops = _sass(k)
        assert sum(o.startswith("MUFU.EX2") for o in ops) >= (5 if tm == "ILi4E" else 9)  # (T + 1) per (row, frame)
        assert any(o.startswith("LDG.E.128") for o in ops), "quads of 4 frames are read as 16-byte words"
    k = _one(resources, "pf_dbeta_fold_kernel")
    assert resources[k]["stack"] == 0


def test_per_frame_pass_holds_both_forward_variants(resources):
    """bv_pass_kernel_pf contains the frozen-parameter consumer (uptake values kept in registers across the phases) and the
    variant that re-evaluates the exponentials at the updated BV parameters of both plateau branches: more than twice the
    special-function instructions of the frozen path alone (72 = 2 quads x 4 frames x 9 exponentials, for T <= 8)"""
    k = _one(resources, "bv_pass_kernel_pf")
    ops = _sass(k)
    assert sum(o.startswith("MUFU.EX2") for o in ops) >= 3 * 72
    assert any(o.startswith("CALL") for o in ops), "derive_spec stays out of line (the tail executes the same instructions)"
