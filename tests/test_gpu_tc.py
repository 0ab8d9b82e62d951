"""tcgen05 / TMEM / TMA GEMM (fira_gemm_bf16_tc) against torch on the same bf16-rounded operands.
bf16 x bf16 products are exact in fp32, so only the fp32 summation order differs: tolerance 1e-4."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).to(DEV).to(torch.bfloat16)


def close(a, b, rtol=1e-4):
    a, b = a.double(), b.double()
    err = (a - b).abs().max().item()
    assert err <= rtol * b.abs().max().item() + 1e-6, f"max err {err:.3e} (ref scale {b.abs().max().item():.3e})"


@pytest.mark.parametrize("M,N,K", [(128, 256, 256), (128, 64, 64), (1000, 256, 256), (41600, 256, 256),
                                   (1920, 1024, 256), (1920, 256, 1024), (300, 72, 200), (257, 3072, 256),
                                   (1920, 24650, 256)])
def test_kmajor_forward(M, N, K):
    from fira_icse_b200 import ops as o
    x, W = rnd(M, K, seed=1), rnd(N, K, seed=2)
    b = torch.randn(N, device=DEV)
    ref = x.float() @ W.float().T
    ldc = (N + 7) // 8 * 8
    c = torch.full((M, ldc), 3.0, device=DEV)
    o.gemm_tc(x, K, 1, W, K, 1, c, ldc, M, N, K)
    close(c[:, :N], ref)
    assert (c[:, N:] == 3.0).all()
    c16 = torch.empty((M, ldc), device=DEV, dtype=torch.bfloat16)
    o.gemm_tc(x, K, 1, W, K, 1, c16, ldc, M, N, K, bias=b, relu=True)
    close(c16[:, :N].float(), torch.relu(ref + b), rtol=1e-2)


def test_rank1_and_splitk():
    from fira_icse_b200 import ops as o
    M, N, K = 700, 256, 512
    x, W = rnd(M, K, seed=1), rnd(N, K, seed=2)
    b, rs, rc = torch.randn(N, device=DEV), torch.randn(M, device=DEV), torch.randn(N, device=DEV)
    ref = x.float() @ W.float().T + b + rs[:, None] * rc[None]
    c = torch.empty((M, N), device=DEV)
    o.gemm_tc(x, K, 1, W, K, 1, c, N, M, N, K, bias=b, rs=rs, rc=rc)
    close(c, ref)
    for s in (2, 3, 8):
        c = torch.full((M, N), 9.0, device=DEV)
        o.gemm_tc(x, K, 1, W, K, 1, c, N, M, N, K, bias=b, rs=rs, rc=rc, splits=s)
        close(c, ref)


@pytest.mark.parametrize("rows,N,K", [(1000, 256, 256), (41600, 256, 256), (1920, 1024, 256), (333, 72, 136),
                                      (1920, 512, 256)])
def test_mnmajor_weight_grad_and_input_grad(rows, N, K):
    """dW[N,K] = dY[rows,N]^T X[rows,K] (both operands MN-major, split over rows) and
    dX[rows,K] = dY[rows,N] W[N,K] (A K-major, B MN-major)."""
    from fira_icse_b200 import ops as o
    dy, x, W = rnd(rows, N, seed=1), rnd(rows, K, seed=2), rnd(N, K, seed=3)
    dW = torch.empty((N, K), device=DEV)
    o.gemm_tc(dy, N, 0, x, K, 0, dW, K, N, K, rows)
    close(dW, dy.float().T @ x.float())
    dW2 = torch.empty((N, K), device=DEV)
    o.gemm_tc(dy, N, 0, x, K, 0, dW2, K, N, K, rows, splits=37)
    close(dW2, dy.float().T @ x.float())
    dx = torch.empty((rows, K), device=DEV, dtype=torch.bfloat16)
    o.gemm_tc(dy, N, 1, W, K, 0, dx, K, rows, K, N)
    close(dx.float(), dy.float() @ W.float(), rtol=1e-2)


@pytest.mark.parametrize("rows,N,K,ld", [(1000, 256, 256, 256), (1920, 1024, 256, 1024), (333, 72, 136, 72), (1920, 768, 256, 768),
                                         (1920, 24650, 256, 24704), (11000, 3072, 256, 3072), (77, 256, 1024, 256)])
def test_weight_grad_with_folded_bias_grad(rows, N, K, ld):
    """fira_gemm_bf16_tc_dbias: dW = dY^T X and db += colsum(dY) from the same launch (dY tiles summed in shared memory)"""
    from fira_icse_b200 import ops as o
    dy = torch.zeros((rows, ld), device=DEV, dtype=torch.bfloat16)
    dy[:, :N] = rnd(rows, N, seed=1)
    x = rnd(rows, K, seed=2)
    pr = o.Prec(True)
    ref_w, ref_b = dy[:, :N].float().T @ x.float(), dy[:, :N].float().sum(0)
    for _ in range(2):                                   # the second pass checks nothing is left behind in the buffers
        dW = torch.full((N, K), 7.0, device=DEV)
        db = torch.zeros(N, device=DEV)
        pr.linear_dw(dy, ld, x, K, rows, N, K, out=dW, dbias=db)
        close(dW, ref_w)
        close(db, ref_b)
