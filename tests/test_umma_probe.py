"""Pin the tcgen05 operand / accumulator layouts on real hardware (B200).

Each case builds the smem byte image of A and B on the host exactly as the production
kernels lay them out (TMA SWIZZLE_128B chunks), issues K/16 tcgen05.mma through
rz_umma_probe and checks the TMEM dump against A @ B^T.
"""
import numpy as np
import pytest
import torch

from tests import umma_layouts as L

pytestmark = pytest.mark.gpu


def _run(M, N, K, a_mn, b_mn, *, lbo_a=None, sbo_a=1024, lbo_b=None, sbo_b=1024, d_off=0, seed=0):
    from radzero_b200 import ops
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((M, K)).astype(np.float16)
    B = rng.standard_normal((N, K)).astype(np.float16)
    a_img = L.image_mn_major(A) if a_mn else L.image_k_major(A)
    b_img = L.image_mn_major(B) if b_mn else L.image_k_major(B)
    if a_mn:
        a_desc = L.smem_desc(0, lbo=(K * 128 if lbo_a is None else lbo_a), sbo=sbo_a)
        a_step = 16 * 128
    else:
        a_desc = L.smem_desc(0, lbo=(0 if lbo_a is None else lbo_a), sbo=sbo_a)
        a_step = 32
    if b_mn:
        b_desc = L.smem_desc(0, lbo=(K * 128 if lbo_b is None else lbo_b), sbo=sbo_b)
        b_step = 16 * 128
    else:
        b_desc = L.smem_desc(0, lbo=(0 if lbo_b is None else lbo_b), sbo=sbo_b)
        b_step = 32
    ncols = max(32, ((N + 7) // 8) * 8)
    out = ops.umma_probe(torch.from_numpy(a_img).cuda(), torch.from_numpy(b_img).cuda(),
                         a_desc, b_desc, a_step, b_step, K // 16, L.idesc_f16(M, N, a_mn, b_mn),
                         d_tmem_offset=d_off, ncols=ncols)
    torch.cuda.synchronize()
    got = out.cpu().numpy()
    want = A.astype(np.float32) @ B.astype(np.float32).T
    return got, want


def _check(got, want, M, lane_shift=0):
    N = want.shape[1]
    err = 0.0
    for r in range(M):
        lane = L.lane_of_row(M, r) + lane_shift
        err = max(err, float(np.abs(got[lane, :N] - want[r]).max()))
    return err


@pytest.mark.parametrize("M,N", [(128, 16), (128, 64), (128, 256), (64, 16), (64, 64)])
def test_k_major_both(M, N):
    got, want = _run(M, N, 64, 0, 0)
    err = _check(got, want, M)
    print(f"K-major M={M} N={N}: max err {err:.3e}")
    assert err < 2e-2


def test_m64_second_half_lanes():
    """Two interleaved M=64 accumulators: the second one addressed at lane offset 16."""
    got, want = _run(64, 16, 64, 0, 0, d_off=(16 << 16))
    err = _check(got, want, 64, lane_shift=16)
    print(f"M=64 at lane offset 16: max err {err:.3e}")
    assert err < 2e-2


@pytest.mark.parametrize("N", [16, 64])
def test_a_mn_major(N):
    """A = K^T read from token-major rows (MN-major), B K-major: the pooling MMA."""
    got, want = _run(128, N, 64, 1, 0)
    err = _check(got, want, 128)
    print(f"A MN-major N={N}: max err {err:.3e}")
    if err >= 2e-2:  # diagnostics: try the alternative LBO/SBO reading
        for lbo, sbo in [(1024, 64 * 128), (64 * 128, 1024), (0, 1024), (128, 1024)]:
            g2, w2 = _run(128, N, 64, 1, 0, lbo_a=lbo, sbo_a=sbo)
            print(f"   alt lbo={lbo} sbo={sbo}: err {_check(g2, w2, 128):.3e}")
    assert err < 2e-2


@pytest.mark.parametrize("N", [64, 256])
def test_b_mn_major(N):
    got, want = _run(128, N, 64, 0, 1)
    err = _check(got, want, 128)
    print(f"B MN-major N={N}: max err {err:.3e}")
    if err >= 2e-2:
        for lbo, sbo in [(1024, 64 * 128), (64 * 128, 1024)]:
            g2, w2 = _run(128, N, 64, 0, 1, lbo_b=lbo, sbo_b=sbo)
            print(f"   alt lbo={lbo} sbo={sbo}: err {_check(g2, w2, 128):.3e}")
    assert err < 2e-2


def test_both_mn_major():
    got, want = _run(128, 128, 64, 1, 1)
    err = _check(got, want, 128)
    print(f"A,B MN-major: max err {err:.3e}")
    assert err < 2e-2
