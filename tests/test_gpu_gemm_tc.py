"""The tcgen05 GEMM (split-bf16 operands, fp32 TMEM accumulation) against the fp32 statement, all three operand
layouts, ragged shapes, strided views, fused epilogues, split-K.  -m gpu."""
import pytest
import torch

from emu_ops import EmuOps
from helpers import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture()
def tc_ops():
    from factorized_b200.cuda_ops import CudaOps, PATH_TC_BF16X3, PATH_SIMT_FP32
    ops = CudaOps()
    old = ops.get_gemm_path()
    ops.set_gemm_path(PATH_TC_BF16X3, min_work=0)
    yield ops
    ops.set_gemm_path(old, min_work=1 << 20)


def g(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed + sum(shape)))


SHAPES = [(128, 16, 16), (128, 128, 32), (256, 256, 64), (1, 1, 1), (5, 3, 7), (70, 130, 33), (640, 128, 300), (300, 400, 128),
          (1000, 24, 5), (129, 257, 65), (96, 512, 464),
          # 16 B-aligned ragged shapes: these take the pipelined cp.async kernel (gemm_tcp.cu)
          (72, 132, 36), (1000, 24, 8), (516, 260, 68), (2048, 2048, 32), (4, 4, 4), (132, 20, 2052), (388, 112, 20),
          # >= 4096 rows: NT / NN pre-split the weight operand into its MMA image (tiles up to 256 columns wide)
          (4100, 400, 128), (4096, 260, 36), (4608, 32, 8), (5000, 513, 72), (4200, 20, 300)]


@pytest.mark.parametrize("mode", ["nt", "nn", "tn"])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_tc_gemm_bf16x3(tc_ops, mode, M, N, K):
    A = g(M, K, seed=1) if mode != "tn" else g(K, M, seed=1)
    B = g(N, K, seed=2) if mode == "nt" else g(K, N, seed=2)
    C0 = g(M, N, seed=3)
    for acc in (False, True):
        c_cpu, c_gpu = C0.clone().double(), C0.clone().cuda()
        EmuOps().gemm(mode, A.double(), B.double(), c_cpu, accumulate=acc)
        tc_ops.gemm(mode, A.cuda(), B.cuda(), c_gpu, accumulate=acc)
        torch.cuda.synchronize()
        assert rel_l2(c_gpu, c_cpu) < 5e-5, (mode, M, N, K, acc, rel_l2(c_gpu, c_cpu))


def test_tc_gemm_split_k_weight_gradient(tc_ops):
    K, M, N = 40960, 128, 400
    A, B = g(K, M, seed=1), g(K, N, seed=2)
    c_cpu, c_gpu = torch.ones(M, N).double(), torch.ones(M, N).cuda()
    EmuOps().gemm("tn", A.double(), B.double(), c_cpu, accumulate=True)
    tc_ops.gemm("tn", A.cuda(), B.cuda(), c_gpu, accumulate=True)
    assert rel_l2(c_gpu, c_cpu) < 5e-5


@pytest.mark.parametrize("TB", [700, 4300])
def test_tc_gemm_views_and_epilogues(tc_ops, TB):
    D = 325
    X = g(TB, D, seed=5)                      # modality slices of x: row pitch 1300 B, not 16 B aligned
    W = g(352, 300, seed=6)
    bias, bias2 = g(352, seed=7), g(352, seed=8)
    rng = torch.tensor([99, 3], dtype=torch.int64)
    for act in (0, 1, 2, 3):
        for drop in (None, (0.4, 5)):
            out_cpu = torch.zeros(TB, 400)
            out_gpu = out_cpu.clone().cuda()
            EmuOps().gemm("nt", X[:, :300], W, out_cpu[:, 10:362], bias=bias, bias2=bias2, act=act, drop=drop, rng=rng)
            tc_ops.gemm("nt", X.cuda()[:, :300], W.cuda(), out_gpu[:, 10:362], bias=bias.cuda(), bias2=bias2.cuda(),
                        act=act, drop=drop, rng=rng.cuda())
            # the same product from a 16 B-aligned copy of the slice (TMA-staged kernel; pre-split weight when TB >= 4096)
            out_al = torch.zeros(TB, 400).cuda()
            tc_ops.gemm("nt", X[:, :300].contiguous().cuda(), W.cuda(), out_al[:, 12:364], bias=bias.cuda(), bias2=bias2.cuda(),
                        act=act, drop=drop, rng=rng.cuda())
            dal = (out_al[:, 12:364] - out_gpu[:, 10:362]).abs()
            assert float((dal > 2e-3).float().mean()) < 1e-4, (act, drop)
            # relu/dropout decisions can flip for |pre-activation| ~ 1e-5; compare with an absolute allowance
            diff = (out_gpu.cpu() - out_cpu).abs()
            assert float((diff > 2e-3).float().mean()) < 1e-4, (act, drop)
            assert float(out_gpu[:, :10].abs().max()) == 0 and float(out_gpu[:, 362:].abs().max()) == 0
    Xa = X[:, 300:305]
    Wa = g(32, 5, seed=9)
    o_cpu, o_gpu = torch.zeros(TB, 32), torch.zeros(TB, 32).cuda()
    EmuOps().gemm("nt", Xa, Wa, o_cpu)
    tc_ops.gemm("nt", X.cuda()[:, 300:305], Wa.cuda(), o_gpu)
    assert rel_l2(o_gpu, o_cpu) < 5e-5
    dY, Wn, Hm = g(TB, 64, seed=9), g(64, 128, seed=10), torch.relu(g(TB, 128, seed=11))
    o_cpu, o_gpu = torch.zeros(TB, 128), torch.zeros(TB, 128).cuda()
    EmuOps().gemm("nn", dY, Wn, o_cpu, mask=Hm, mask_scale=2.0)
    tc_ops.gemm("nn", dY.cuda(), Wn.cuda(), o_gpu, mask=Hm.cuda(), mask_scale=2.0)
    assert rel_l2(o_gpu, o_cpu) < 5e-5


def test_tc_gemm_plain_bf16_path():
    from factorized_b200.cuda_ops import CudaOps, PATH_TC_BF16
    ops = CudaOps()
    old = ops.get_gemm_path()
    try:
        ops.set_gemm_path(PATH_TC_BF16, min_work=0)
        A, B = g(300, 200, seed=1), g(150, 200, seed=2)
        c_cpu, c_gpu = torch.zeros(300, 150), torch.zeros(300, 150).cuda()
        EmuOps().gemm("nt", A, B, c_cpu)
        ops.gemm("nt", A.cuda(), B.cuda(), c_gpu)
        e = rel_l2(c_gpu, c_cpu)
        assert 1e-4 < e < 1e-2, e           # one bf16 pass: ~2^-9 per operand
    finally:
        ops.set_gemm_path(old, min_work=1 << 20)


@pytest.mark.parametrize("K,M,N", [(40960, 128, 400), (700, 352, 300), (5000, 96, 255), (640, 416, 104), (300, 16, 16)])
def test_tc_gemm_fused_bias_gradient(tc_ops, K, M, N):
    """TN GEMM with colsum_out: dW += dY^T X and db += colsum(dY) in one launch (ones column of X)."""
    dY, X = g(K, M, seed=1), g(K, N, seed=2)
    w_cpu, w_gpu = torch.ones(M, N).double(), torch.ones(M, N).cuda()
    b_cpu, b_gpu = torch.ones(M).double(), torch.ones(M).cuda()
    EmuOps().gemm("tn", dY.double(), X.double(), w_cpu, accumulate=True, colsum_out=b_cpu)
    tc_ops.gemm("tn", dY.cuda(), X.cuda(), w_gpu, accumulate=True, colsum_out=b_gpu)
    assert rel_l2(w_gpu, w_cpu) < 5e-5 and rel_l2(b_gpu, b_cpu) < 5e-5, (rel_l2(w_gpu, w_cpu), rel_l2(b_gpu, b_cpu))


@pytest.mark.parametrize("K,M,N1,N2", [(40960, 352, 300, 88), (4096, 32, 5, 8), (700, 128, 400, 64), (5000, 96, 20, 80),
                                        (300, 16, 16, 16)])
def test_tc_gemm_tn_pair(tc_ops, K, M, N1, N2):
    """Two weight gradients that share dY in one launch: C1 += dY^T X1 (+ bias gradient), C2 += dY^T X2."""
    ld1 = (N1 + 3) // 4 * 4                                    # x slices live in 16 B-aligned padded buffers
    dY, X1p, X2 = g(K, M, seed=1), g(K, ld1, seed=2), g(K, N2, seed=3)
    X1 = X1p[:, :N1]
    w1_cpu, w2_cpu, b_cpu = torch.ones(M, N1).double(), torch.ones(M, N2).double(), torch.ones(M).double()
    w1_gpu, w2_gpu, b_gpu = torch.ones(M, N1).cuda(), torch.ones(M, N2).cuda(), torch.ones(M).cuda()
    EmuOps().gemm_tn_pair(dY.double(), X1.double(), w1_cpu, b_cpu, X2.double(), w2_cpu)
    tc_ops.gemm_tn_pair(dY.cuda(), X1p.cuda()[:, :N1], w1_gpu, b_gpu, X2.cuda(), w2_gpu)
    torch.cuda.synchronize()
    e = (rel_l2(w1_gpu, w1_cpu), rel_l2(w2_gpu, w2_cpu), rel_l2(b_gpu, b_cpu))
    assert max(e) < 5e-5, e
