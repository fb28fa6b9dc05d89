"""tcgen05 GEMMs of the batched mode (csrc/bg_gemm.cu) against fp64 matmuls on the same operands.

gemm1: H[f,b] = sum_{a,r} N_a[r,f] c_a[r,b];  gemm2: Acc[a*Rp+r, b] = sum_f N_a[r,f] e[f,b] (summed over the k-split partials).
The fp32 replica-side operand is split into NS bf16 pieces; tolerance: NS=3 keeps ~24 mantissa bits (rel 6e-6 of the
column scale, fp32 accumulation over K), NS=2 ~16 bits (1e-4), NS=1 plain bf16 (2e-2)."""
import numpy as np
import pytest
import torch

from jaxent_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _stream():
    return int(torch.cuda.current_stream().cuda_stream)


def split_bf16(x: torch.Tensor, ns: int):
    """fp32 -> ns bf16 pieces with x ~= sum(pieces) (round-to-nearest at every level)."""
    out, rem = [], x.float().clone()
    for _ in range(ns):
        p = rem.to(torch.bfloat16)
        out.append(p)
        rem = rem - p.float()
    return out


def pack_cc3(c: torch.Tensor, Rp: int, Bn: int, ns: int):
    """c: (2, R, B) fp32 -> [NS][2][Rp/8][Bn][8] bf16"""
    _, R, B = c.shape
    full = torch.zeros((2, Rp, Bn), dtype=torch.float32, device=c.device)
    full[:, :R, :B] = c
    pieces = split_bf16(full, ns)
    t = torch.stack(pieces)  # (NS, 2, Rp, Bn)
    return t.view(ns, 2, Rp // 8, 8, Bn).permute(0, 1, 2, 4, 3).contiguous()


def pack_e3(e: torch.Tensor, Fp: int, Bn: int, ns: int):
    """e: (F, B) fp32 -> [Fp/128][NS][16][Bn][8] bf16"""
    F, B = e.shape
    full = torch.zeros((Fp, Bn), dtype=torch.float32, device=e.device)
    full[:F, :B] = e
    pieces = split_bf16(full, ns)
    t = torch.stack(pieces)  # (NS, Fp, Bn)
    return t.view(ns, Fp // 128, 16, 8, Bn).permute(1, 0, 2, 4, 3).contiguous()


def _features(R, F, seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    heavy = torch.randint(0, 60, (R, F), device="cuda", generator=g, dtype=torch.int32).float()
    acc = torch.randint(0, 4, (R, F), device="cuda", generator=g, dtype=torch.int32).float()
    return heavy, acc


def _pack(lib, heavy, acc):
    R, F = heavy.shape
    nbytes = lib.jaxent_bg_features_bytes(R, F)
    out = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    flag = torch.ones(1, dtype=torch.int32, device="cuda")
    L.check(lib.jaxent_bg_pack_features(heavy.data_ptr(), acc.data_ptr(), R, F, F, out.data_ptr(), flag.data_ptr(), _stream()))
    assert int(flag.item()) == 1
    return out


@pytest.mark.parametrize("R,F,B,ns", [(53, 500, 16, 3), (300, 1000, 64, 3), (500, 3000, 256, 3), (500, 3000, 256, 2), (130, 777, 48, 1)])
def test_gemm1_matches_fp64(R, F, B, ns):
    lib = L.load()
    heavy, acc = _features(R, F, 1)
    feat = _pack(lib, heavy, acc)
    Rp, Fp, Bn = (R + 127) // 128 * 128, (F + 127) // 128 * 128, (B + 15) // 16 * 16
    g = torch.Generator(device="cuda")
    g.manual_seed(2)
    c = torch.randn((2, R, B), device="cuda", generator=g) * 1e-3
    cc3 = pack_cc3(c, Rp, Bn, ns)
    H = torch.full((Fp, Bn), float("nan"), device="cuda")
    L.check(lib.jaxent_bg_gemm1(feat.data_ptr(), cc3.data_ptr(), H.data_ptr(), Fp // 128, Rp // 8, Bn, ns, 148, _stream()))
    torch.cuda.synchronize()
    ref = heavy.double().T @ c[0].double() + acc.double().T @ c[1].double()  # (F, B)
    got = H[:F, :B].double()
    scale = ref.abs().max()
    tol = {3: 6e-6, 2: 1e-4, 1: 2e-2}[ns]
    assert torch.isfinite(H).all()
    assert float((got - ref).abs().max() / scale) < tol
    assert float(H[F:, :].abs().max()) == 0.0 if Fp > F else True


@pytest.mark.parametrize("R,F,B,ns,ksplit", [(53, 500, 16, 3, 1), (300, 1000, 64, 3, 3), (500, 3000, 256, 3, 5), (500, 3000, 256, 2, 18), (130, 777, 48, 1, 2)])
def test_gemm2_matches_fp64(R, F, B, ns, ksplit):
    lib = L.load()
    heavy, acc = _features(R, F, 3)
    feat = _pack(lib, heavy, acc)
    Rp, Fp, Bn = (R + 127) // 128 * 128, (F + 127) // 128 * 128, (B + 15) // 16 * 16
    g = torch.Generator(device="cuda")
    g.manual_seed(4)
    e = torch.rand((F, B), device="cuda", generator=g) ** 4
    e3 = pack_e3(e, Fp, Bn, ns)
    P = torch.full((ksplit, 2 * Rp, Bn), float("nan"), device="cuda")
    L.check(lib.jaxent_bg_gemm2(feat.data_ptr(), e3.data_ptr(), P.data_ptr(), Fp // 128, Rp // 8, Bn, ns, ksplit, _stream()))
    torch.cuda.synchronize()
    assert torch.isfinite(P).all()
    got = P.double().sum(0)
    ref_h = heavy.double() @ e.double()
    ref_a = acc.double() @ e.double()
    tol = {3: 6e-6, 2: 1e-4, 1: 2e-2}[ns]
    for arr, ref in ((0, ref_h), (1, ref_a)):
        blk = got[arr * Rp : arr * Rp + R, :B]
        assert float((blk - ref).abs().max() / ref.abs().max()) < tol
        assert float(got[arr * Rp + R : (arr + 1) * Rp].abs().max()) == 0.0 if Rp > R else True
