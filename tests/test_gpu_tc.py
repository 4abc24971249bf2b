"""Bring-up / regression tests of the tcgen05 machinery (TMA tensor maps, UMMA smem and
instruction descriptors, TMEM accumulators, the warp-specialised skeleton) through the
pvs_debug_tc_gemm validation export, against torch fp64 matmul."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def tf32_round(x):
    """round-to-nearest-even to 10 mantissa bits (torch, on device)"""
    u = x.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    u = (u + 0xFFF + ((u >> 13) & 1)) & ~0x1FFF
    u = torch.where(u >= 2 ** 31, u - 2 ** 32, u)
    return u.to(torch.int32).view(torch.float32)


def split(x):
    hi = tf32_round(x)
    return hi, tf32_round(x - hi)


@pytest.fixture(scope="module")
def nat():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from pyvisim_b200 import _native
    return _native


def run(nat, mode, a_hi, a_lo, b_hi, b_lo, m, n, k, block_n):
    c = torch.full((m, n), float("nan"), dtype=torch.float32, device="cuda")
    p = lambda t: None if t is None else t.data_ptr()
    nat.check(nat.lib().pvs_debug_tc_gemm(mode, p(a_hi), p(a_lo), p(b_hi), p(b_lo), c.data_ptr(), m, n, k, block_n, None))
    torch.cuda.synchronize()
    return c


@pytest.mark.parametrize("block_n", [64, 128, 256])
@pytest.mark.parametrize("m,n,k", [(128, 256, 32), (128, 256, 128), (300, 512, 96), (1000, 256, 512)])
def test_kmajor_tf32_and_3xtf32(nat, m, n, k, block_n):
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = torch.randn((m, k), device="cuda", generator=g)
    b = torch.randn((n, k), device="cuda", generator=g)
    ref = (a.double() @ b.double().T)
    a_hi, a_lo = split(a)
    b_hi, b_lo = split(b)
    c1 = run(nat, 0, a_hi, None, b_hi, None, m, n, k, block_n)
    ref1 = a_hi.double() @ b_hi.double().T
    assert torch.isfinite(c1).all()
    assert (c1.double() - ref1).abs().max().item() <= 1e-4 * k ** 0.5 + 1e-5, "tf32 single pass vs tf32-exact reference"
    c3 = run(nat, 1, a_hi, a_lo, b_hi, b_lo, m, n, k, block_n)
    # the tensor core's fp32 accumulation truncates: error grows ~linearly with the number
    # of accumulation steps (measured 4.8e-4 at k=512 for N(0,1) operands)
    err = (c3.double() - ref).abs().max().item()
    assert err <= 2e-6 * k + 5e-5, f"3xTF32 max abs err {err}"


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (257, 384, 1024)])
def test_kmajor_bf16(nat, m, n, k):
    g = torch.Generator(device="cuda").manual_seed(7)
    a = torch.randn((m, k), device="cuda", generator=g).bfloat16()
    b = torch.randn((n, k), device="cuda", generator=g).bfloat16()
    c = run(nat, 2, a, None, b, None, m, n, k, 128)
    ref = a.double() @ b.double().T
    assert (c.double() - ref).abs().max().item() <= 1e-3 * k ** 0.5


@pytest.mark.parametrize("block_n", [128, 256])
@pytest.mark.parametrize("m,n,k", [(128, 256, 32), (128, 256, 200), (128, 512, 2000), (256, 256, 77)])
def test_mnmajor_tf32_and_3xtf32(nat, m, n, k, block_n):
    """C = A^T B with A [k,m], B [k,n]; k not a multiple of the 32-row stage exercises the
    TMA out-of-bounds zero fill on the contraction rows."""
    g = torch.Generator(device="cuda").manual_seed(k)
    a = torch.randn((k, m), device="cuda", generator=g)
    b = torch.randn((k, n), device="cuda", generator=g)
    ref = a.double().T @ b.double()
    a_hi, a_lo = split(a)
    b_hi, b_lo = split(b)
    c1 = run(nat, 3, a_hi, None, b_hi, None, m, n, k, block_n)
    ref1 = a_hi.double().T @ b_hi.double()
    assert (c1.double() - ref1).abs().max().item() <= 1e-4 * k ** 0.5 + 1e-5
    c3 = run(nat, 4, a_hi, a_lo, b_hi, b_lo, m, n, k, block_n)
    assert (c3.double() - ref).abs().max().item() <= 2e-6 * k + 5e-5


@pytest.mark.parametrize("block_n", [128, 256])
@pytest.mark.parametrize("m,n,k", [(256, 256, 64), (256, 512, 1024), (700, 300, 256), (1000, 1000, 2048)])
def test_pair_bf16(nat, m, n, k, block_n):
    """cta_group::2: two CTAs of a cluster issue one MMA; each loads half of A and half of B."""
    g = torch.Generator(device="cuda").manual_seed(11 + m)
    a = torch.randn((m, k), device="cuda", generator=g).bfloat16()
    b = torch.randn((n, k), device="cuda", generator=g).bfloat16()
    c = run(nat, 5, a, None, b, None, m, n, k, block_n)
    ref = a.double() @ b.double().T
    assert torch.isfinite(c).all()
    assert (c.double() - ref).abs().max().item() <= 1e-3 * k ** 0.5


@pytest.mark.parametrize("block_n", [128, 256])
@pytest.mark.parametrize("m,n,k", [(256, 256, 32), (300, 512, 96), (1000, 256, 512)])
def test_pair_3xtf32(nat, m, n, k, block_n):
    g = torch.Generator(device="cuda").manual_seed(m + n + k)
    a = torch.randn((m, k), device="cuda", generator=g)
    b = torch.randn((n, k), device="cuda", generator=g)
    ref = (a.double() @ b.double().T)
    a_hi, a_lo = split(a)
    b_hi, b_lo = split(b)
    c3 = run(nat, 6, a_hi, a_lo, b_hi, b_lo, m, n, k, block_n)
    err = (c3.double() - ref).abs().max().item()
    assert err <= 2e-6 * k + 5e-5, f"pair 3xTF32 max abs err {err}"


@pytest.mark.parametrize("m,n", [(256, 256), (5000, 256), (130, 200)])
def test_pair_3xtf32_resident_b(nat, m, n):
    """weights-resident variant: all k-blocks of B loaded once per kernel, reused by every tile"""
    k = 128
    g = torch.Generator(device="cuda").manual_seed(m)
    a = torch.randn((m, k), device="cuda", generator=g)
    b = torch.randn((n, k), device="cuda", generator=g)
    ref = (a.double() @ b.double().T)
    a_hi, a_lo = split(a)
    b_hi, b_lo = split(b)
    c3 = run(nat, 7, a_hi, a_lo, b_hi, b_lo, m, n, k, 256)
    err = (c3.double() - ref).abs().max().item()
    assert err <= 2e-6 * k + 5e-5, f"pair resident-B 3xTF32 max abs err {err}"
