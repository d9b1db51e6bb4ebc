"""Each CUDA primitive (through the C ABI / ctypes binding) against its torch statement in emu_ops.py on
seeded random inputs, including ragged / awkward sizes.  Needs a B200: run with -m gpu."""
import pytest
import torch

from emu_ops import EmuOps, keep_mask
from helpers import rel_l2

pytestmark = pytest.mark.gpu

TOL = 2e-5


@pytest.fixture(scope="module")
def ops():
    from factorized_b200.cuda_ops import CudaOps
    return CudaOps()


def g(*shape, seed=0, scale=1.0):
    gen = torch.Generator().manual_seed(seed + sum(shape))
    return torch.randn(*shape, generator=gen) * scale


def both(fn_name, cpu_args, ops, emu=None, **kw):
    """run emu on cpu tensors and cuda op on device copies; returns (cpu_args, gpu_args)."""
    emu = emu or EmuOps()
    dev = [a.cuda() if isinstance(a, torch.Tensor) else a for a in cpu_args]
    getattr(emu, fn_name)(*cpu_args, **kw)
    kw2 = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
    getattr(ops, fn_name)(*dev, **kw2)
    torch.cuda.synchronize()
    return cpu_args, dev


@pytest.mark.parametrize("mode", ["nt", "nn", "tn"])
@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (5, 3, 7), (64, 64, 16), (70, 130, 33), (640, 128, 300), (37, 400, 128),
                                   (12, 20, 5000)])
def test_gemm_plain(ops, mode, M, N, K):
    A = g(M, K, seed=1) if mode != "tn" else g(K, M, seed=1)
    B = g(N, K, seed=2) if mode == "nt" else g(K, N, seed=2)
    C0 = g(M, N, seed=3)
    for acc in (False, True):
        if mode == "tn" and not acc and K >= 1024:
            pass
        c_cpu, c_gpu = C0.clone(), C0.clone().cuda()
        EmuOps().gemm(mode, A, B, c_cpu, accumulate=acc)
        ops.gemm(mode, A.cuda(), B.cuda(), c_gpu, accumulate=acc)
        assert rel_l2(c_gpu, c_cpu) < TOL, (mode, M, N, K, acc)


def test_gemm_strided_views_and_epilogues(ops):
    T_B, D = 96, 41
    X = g(T_B, D, seed=5)
    W = g(24, 30, seed=6)            # use a column slice of both
    bias, bias2 = g(24, seed=7), g(24, seed=8)
    rng = torch.tensor([99, 3], dtype=torch.int64)
    for act in (0, 1, 2, 3):
        for drop in (None, (0.4, 5)):
            out_cpu = torch.zeros(T_B, 50)
            out_gpu = out_cpu.clone().cuda()
            EmuOps().gemm("nt", X[:, 7:27], W[:, 3:23], out_cpu[:, 10:34], bias=bias, bias2=bias2, act=act, drop=drop, rng=rng)
            ops.gemm("nt", X.cuda()[:, 7:27], W.cuda()[:, 3:23], out_gpu[:, 10:34], bias=bias.cuda(), bias2=bias2.cuda(),
                     act=act, drop=drop, rng=rng.cuda())
            assert rel_l2(out_gpu, out_cpu) < TOL, (act, drop)
            assert float(out_gpu[:, :10].abs().max()) == 0 and float(out_gpu[:, 34:].abs().max()) == 0
    # relu/dropout mask epilogue of the data-gradient GEMM
    dY, Wn, Hm = g(T_B, 24, seed=9), g(24, 17, seed=10), torch.relu(g(T_B, 17, seed=11))
    o_cpu, o_gpu = torch.zeros(T_B, 17), torch.zeros(T_B, 17).cuda()
    EmuOps().gemm("nn", dY, Wn, o_cpu, mask=Hm, mask_scale=2.0)
    ops.gemm("nn", dY.cuda(), Wn.cuda(), o_gpu, mask=Hm.cuda(), mask_scale=2.0)
    assert rel_l2(o_gpu, o_cpu) < TOL


@pytest.mark.parametrize("mode,M,N,K,act,drop,persistent", [
    ("nt", 40960, 400, 128, 0, None, True),          # three N tiles (160 / 160 / 96)
    ("nt", 8192, 128, 400, 1, None, True),           # streaming mode (the image does not fit beside the ring)
    ("nn", 12288, 400, 384, 0, None, True),          # NN: MN-major weight image
    ("nt", 4100, 36, 20, 2, None, True),             # ragged last row tile, N tail clipped by the TMA store, K < one stage
    ("nn", 5000, 300, 104, 3, None, True),           # K tail (104 = 3 stages + 8)
    ("nt", 8192, 64, 128, 2, (0.3, 5), True),        # dropout in the epilogue
    ("nt", 6000, 480, 300, 1, (0.5, 2), True),       # three streaming N tiles, odd tile count per CTA (unpaired last tile)
    ("nt", 4096, 34, 64, 0, None, False),            # N not a multiple of 4: the per-tile kernel serves it
    ("nt", 4000, 128, 128, 0, None, False),          # M < 4096: no weight-image workspace, per-tile kernel
])
@pytest.mark.parametrize("residency", [0, 1])
def test_gemm_persistent_kernel(ops, mode, M, N, K, act, drop, persistent, residency):
    """The persistent streamed GEMM (csrc/gemm_ps.cu: TMA-fed ring, two MMA issuers, TMA-store epilogue) on the shapes /
    epilogues it accepts, against fp64, with the weight image streamed (the default) and resident in shared memory where it
    fits; the launch counter proves which kernel served the call; the padding columns of a wider output buffer must stay
    untouched."""
    ops.lib.mfm_debug_gemm_ps_residency(residency)
    try:
        _gemm_persistent_case(ops, mode, M, N, K, act, drop, persistent)
    finally:
        ops.lib.mfm_debug_gemm_ps_residency(0)


def _gemm_persistent_case(ops, mode, M, N, K, act, drop, persistent):
    gen = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=gen).cuda()
    B = (torch.randn((N, K) if mode == "nt" else (K, N), generator=gen) / K ** 0.5).cuda()
    ldc = (N + 3) // 4 * 4 + 4
    Cfull = torch.full((M, ldc), 7.0).cuda()
    C = Cfull[:, :N]
    bias = torch.randn(N, generator=gen).cuda()
    rng = torch.tensor([1234, 3], dtype=torch.int64)
    n0 = ops.lib.mfm_debug_gemm_ps_count()
    ops.gemm(mode, A, B, C, bias=bias, act=act, drop=drop, rng=rng.cuda())
    torch.cuda.synchronize()
    assert ops.lib.mfm_debug_gemm_ps_count() - n0 == (1 if persistent else 0)
    ref = A.double() @ (B.double().t() if mode == "nt" else B.double()) + bias.double()
    ref = [ref, ref.clamp_min(0), torch.tanh(ref), torch.sigmoid(ref)][act]
    if drop:
        ref = ref * keep_mask(rng, drop[1], drop[0], M, N).cuda().double() / (1 - drop[0])
    assert rel_l2(C.double(), ref) < TOL
    assert bool((Cfull[:, N:] == 7.0).all())


@pytest.mark.parametrize("mode,M,N,K,kind", [("nn", 40960, 128, 400, "mask"), ("nn", 8192, 400, 128, "acc"), ("nt", 4100, 64, 72, "acc"),
                                             ("nn", 6000, 128, 64, "mask"), ("nt", 5000, 36, 20, "mask")])
def test_gemm_persistent_kernel_operand_epilogues(ops, mode, M, N, K, kind):
    """The data-gradient epilogues of the persistent GEMM: ReLU mask (dA = (dY W) * (A > 0) * scale) and accumulate (C += ...),
    their operand tile TMA-loaded one block ahead; against fp64, launch counter checked, padding untouched."""
    gen = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=gen).cuda()
    B = (torch.randn((N, K) if mode == "nt" else (K, N), generator=gen) / K ** 0.5).cuda()
    ldc = N + 8
    Cfull = torch.randn(M, ldc, generator=gen).cuda()
    C0 = Cfull.clone()
    C = Cfull[:, :N]
    Hfull = torch.relu(torch.randn(M, N + 4, generator=gen)).cuda()
    n0 = ops.lib.mfm_debug_gemm_ps_count()
    ref = A.double() @ (B.double().t() if mode == "nt" else B.double())
    if kind == "mask":
        ops.gemm(mode, A, B, C, mask=Hfull[:, :N], mask_scale=2.0)
        ref = torch.where(Hfull[:, :N] > 0, ref * 2.0, torch.zeros_like(ref))
    else:
        ops.gemm(mode, A, B, C, accumulate=True)
        ref = ref + C0[:, :N].double()
    torch.cuda.synchronize()
    assert ops.lib.mfm_debug_gemm_ps_count() - n0 == 1
    assert rel_l2(C.double(), ref) < TOL
    assert torch.equal(Cfull[:, N:], C0[:, N:])


@pytest.mark.parametrize("path", ["simt_fp32", "tcgen05_bf16x3"])
@pytest.mark.parametrize("M,N,K,ldx,want_xhat", [(640, 300, 104, 300, False), (4500, 300, 104, 300, True), (700, 5, 24, 8, True),
                                                 (4500, 300, 104, 300, False), (8200, 20, 24, 24, False),
                                                   (333, 20, 24, 20, False), (64, 7, 9, 7, True)])
def test_gemm_mse_fused_reconstruction_head(ops, path, M, N, K, ldx, want_xhat):
    """mfm_gemm_mse: x_hat = A W^T + b with the MSE term and its gradient produced in the GEMM epilogue
    (mfm_model.py:88-90 + mfm_mosi.py:437); x_hat is written only on request."""
    from factorized_b200.cuda_ops import PATH_SIMT_FP32, PATH_TC_BF16X3
    old = ops.get_gemm_path()
    ops.set_gemm_path(PATH_SIMT_FP32 if path == "simt_fp32" else PATH_TC_BF16X3)
    try:
        A, W, b = g(M, K, seed=1), g(N, K, seed=2, scale=0.2), g(N, seed=3)
        xfull = g(M, ldx, seed=4)
        x = xfull[:, :N]
        slot_c, slot_g = torch.tensor([0.25]), torch.tensor([0.25]).cuda()
        d_c, d_g = torch.zeros(M, N), torch.zeros(M, N).cuda()
        h_c = torch.zeros(M, N) if want_xhat else None
        h_g = torch.zeros(M, N).cuda() if want_xhat else None
        ls, gs = 1.0 / (M * N), 2.0 * 0.5 / (M * N)
        EmuOps().gemm_mse(A, W, b, x, ls, gs, slot_c, d_c, h_c)
        ops.gemm_mse(A.cuda(), W.cuda(), b.cuda(), xfull.cuda()[:, :N], ls, gs, slot_g, d_g, h_g)
        torch.cuda.synchronize()
        assert abs(float(slot_g) - float(slot_c)) < 1e-4 * abs(float(slot_c)), (float(slot_g), float(slot_c))
        assert rel_l2(d_g, d_c) < 1e-4
        if want_xhat:
            assert rel_l2(h_g, h_c) < 1e-4
    finally:
        ops.set_gemm_path(old)


def _lstm_case(T, B, h, gx_steps, seed, ld_extra=0):
    W = g(4 * h, h, seed=seed, scale=0.3)
    gx = g(gx_steps * B, 4 * h, seed=seed + 1)
    bias_rest = g(4 * h, seed=seed + 2) if gx_steps < T else None
    hs = torch.full(((T + 1) * B, h + ld_extra), 7.0)
    cs = torch.full(((T + 1) * B, h + ld_extra), 7.0)
    gates = torch.zeros(T * B, 4 * h)
    return dict(T=T, B=B, h=h, gx=gx, gx_steps=gx_steps, bias_rest=bias_rest, W=W,
                hs=hs[:, ld_extra:], cs=cs[:, ld_extra:], gates=gates)


def _to_dev(c):
    return {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in c.items()}


def _dev_view(c_cpu, key, base_cpu, base_gpu):
    return None


def _variant_counts(ops):
    return [int(ops.lib.mfm_debug_lstm_variant_count(i)) for i in range(6)]


@pytest.mark.parametrize("chains", [1, 2])
def test_lstm_one_and_two_chains_per_cta(ops, chains):
    """Both CTA shapes of the recurrence kernels (one / two batch sub-tiles per CTA) on the same cells."""
    cells = [_lstm_case(5, 150, h, gs, seed=h) for h, gs in ((88, 5), (24, 1), (64, 5), (104, 1))]
    cbs = [dict(T=5, B=150, h=c["h"], gates=c["gates"], cs=c["cs"], W=c["W"], dh_all=g(5 * 150, c["h"], seed=3), dh_last=None,
                dc_ext=g(5 * 150, c["h"], seed=4), dG=torch.zeros(5 * 150, 4 * c["h"]), dc_scratch=torch.zeros(150, c["h"]))
           for c in cells]
    EmuOps().lstm_fwd(cells)
    EmuOps().lstm_bwd(cbs)
    dev, dbs = [_to_dev(c) for c in cells], [_to_dev(c) for c in cbs]
    assert ops.lib.mfm_debug_lstm_force_chains(chains) == 0
    try:
        n0 = [int(ops.lib.mfm_debug_lstm_variant_count(i)) for i in (6, 7)]
        ops.lstm_fwd(dev)
        ops.lstm_bwd(dbs)
        torch.cuda.synchronize()
        n1 = [int(ops.lib.mfm_debug_lstm_variant_count(i)) for i in (6, 7)]
    finally:
        ops.lib.mfm_debug_lstm_force_chains(0)
    assert [a - b for a, b in zip(n1, n0)] == ([2, 0] if chains == 1 else [0, 2])
    for c, d in zip(cells, dev):
        assert rel_l2(d["hs"], c["hs"]) < 1e-4 and rel_l2(d["gates"], c["gates"]) < 1e-4, c["h"]
    for c, d in zip(cbs, dbs):
        assert rel_l2(d["dG"], c["dG"]) < 1e-4, c["h"]


# (T, B, h, gx_steps, ld_extra): every layout of the recurrence kernels -- replicated unit rows (h <= 32: x4, <= 64: x2),
# three and four lane quadrants, 16-row chains (h = 104 backward), the CUDA-core kernel (h > 128) -- on ragged batches,
# plus the production sizes: many CTAs of full 64-row tiles with a ragged last one (B = 2405), decoder-like gx_steps = 1
@pytest.mark.parametrize("T,B,h,gx_steps,ld_extra", [(3, 5, 6, 3, 0), (4, 19, 32, 4, 3), (5, 33, 104, 1, 0), (2, 9, 8, 2, 0),
                                                     (3, 17, 128, 3, 0), (2, 6, 300, 1, 0), (20, 70, 88, 20, 112),
                                                     (3, 100, 64, 3, 0), (3, 70, 48, 3, 5), (2, 45, 96, 2, 0), (4, 77, 80, 4, 0),
                                                     (3, 130, 24, 1, 0), (2, 67, 108, 2, 0),
                                                     (4, 2405, 88, 4, 0), (3, 2048, 104, 1, 0), (3, 2100, 32, 3, 0)])
def test_lstm_fwd_bwd(ops, T, B, h, gx_steps, ld_extra):
    c = _lstm_case(T, B, h, gx_steps, seed=T + B + h, ld_extra=ld_extra)
    # device copies that preserve the strided views
    hs_full = torch.full(((T + 1) * B, h + ld_extra), 7.0).cuda()
    cs_full = torch.full(((T + 1) * B, h + ld_extra), 7.0).cuda()
    cg = _to_dev({k: v for k, v in c.items() if k not in ("hs", "cs")})
    cg["hs"], cg["cs"] = hs_full[:, ld_extra:], cs_full[:, ld_extra:]
    EmuOps().lstm_fwd([c])
    ops.lstm_fwd([cg])
    torch.cuda.synchronize()
    for k in ("hs", "cs", "gates"):
        assert rel_l2(cg[k], c[k]) < 1e-4, (k, T, B, h, rel_l2(cg[k], c[k]))      # tensor-core recurrence: bf16x3
    if ld_extra:
        assert float((hs_full[:, :ld_extra] - 7.0).abs().max()) == 0.0      # neighbours untouched
    # backward on the emulator's forward state
    for variant in range(3):
        dh_all = g(T * B, h, seed=11) if variant in (0, 2) else None
        dh_last = g(B, h, seed=12) if variant in (1, 2) else None
        dc_ext = g(T * B, h, seed=13) if variant == 2 else None
        cb = dict(T=T, B=B, h=h, gates=c["gates"], cs=c["cs"], W=c["W"], dh_all=dh_all, dh_last=dh_last, dc_ext=dc_ext,
                  dG=torch.zeros(T * B, 4 * h), dc_scratch=torch.zeros(B, h))
        cbg = _to_dev({k: v for k, v in cb.items() if k != "cs"})
        cs_dev = torch.zeros((T + 1) * B, h + ld_extra).cuda()
        cs_dev[:, ld_extra:] = c["cs"].cuda()
        cbg["cs"] = cs_dev[:, ld_extra:]
        EmuOps().lstm_bwd([cb])
        ops.lstm_bwd([cbg])
        torch.cuda.synchronize()
        assert rel_l2(cbg["dG"], cb["dG"]) < 1e-4, (variant, T, B, h, rel_l2(cbg["dG"], cb["dG"]))


@pytest.mark.parametrize("T,B,h", [(5, 40, 24), (3, 33, 88), (1, 8, 16)])
def test_lstm_duplicate_cell_history_and_second_cell_gradient(ops, T, B, h):
    """cs_dup (forward): every c block is written twice, here into the cat(c_{t-1}, c_t) layout the MFN attention reads;
    dc_ext2 (backward): the second external cell gradient, added for t < T-1."""
    c = _lstm_case(T, B, h, T, seed=3 * T + B + h)
    cg = _to_dev({k: v for k, v in c.items() if k not in ("hs", "cs")})
    cg["hs"] = torch.zeros((T + 1) * B, h).cuda()
    CS2 = torch.full(((T + 2) * B, 2 * h), 7.0).cuda()
    cg["cs"], cg["cs_dup"] = CS2[:(T + 1) * B, h:], CS2[B:, :h]
    EmuOps().lstm_fwd([c])
    ops.lstm_fwd([cg])
    torch.cuda.synchronize()
    assert rel_l2(cg["cs"], c["cs"]) < 1e-4 and rel_l2(cg["cs_dup"], c["cs"]) < 1e-4
    cstar = CS2[B:(T + 1) * B].cpu()                              # row block t = [c_{t-1} | c_t]
    assert rel_l2(cstar[:, :h], c["cs"][:T * B]) < 1e-4 and rel_l2(cstar[:, h:], c["cs"][B:]) < 1e-4
    dcs = g(T * B, 2 * h, seed=21)                                  # gradient of that concatenation
    cb = dict(T=T, B=B, h=h, gates=c["gates"], cs=c["cs"], W=c["W"], dh_all=None, dh_last=g(B, h, seed=12),
              dc_ext=dcs[:, h:], dc_ext2=(dcs[B:, :h] if T > 1 else None), dG=torch.zeros(T * B, 4 * h),
              dc_scratch=torch.zeros(B, h))
    dcs_dev = dcs.cuda()
    cbg = _to_dev({k: v for k, v in cb.items() if k not in ("dc_ext", "dc_ext2")})
    cbg["dc_ext"], cbg["dc_ext2"] = dcs_dev[:, h:], (dcs_dev[B:, :h] if T > 1 else None)
    EmuOps().lstm_bwd([cb])
    ops.lstm_bwd([cbg])
    torch.cuda.synchronize()
    assert rel_l2(cbg["dG"], cb["dG"]) < 1e-4


@pytest.mark.parametrize("B,hs", [(40, ((32, 6), (8, 6), (80, 6), (24, 1))),
                                  (448, ((32, 6), (8, 6), (80, 6), (88, 6), (64, 6), (48, 6)))])     # the step's 6-cell launch
def test_lstm_multi_cell_launch(ops, B, hs):
    cells = [_lstm_case(6, B, h, gs, seed=h) for h, gs in hs]
    dev = [_to_dev(c) for c in cells]
    EmuOps().lstm_fwd(cells)
    before = _variant_counts(ops)
    ops.lstm_fwd(dev)
    torch.cuda.synchronize()
    after = _variant_counts(ops)
    assert (after[0] + after[1]) - (before[0] + before[1]) == len(cells), (before, after)   # every cell on the tensor-core chains
    for c, d in zip(cells, dev):
        for k in ("hs", "cs", "gates"):
            assert rel_l2(d[k], c[k]) < 1e-4, (k, c["h"], rel_l2(d[k], c[k]))
    # backward of all cells in one launch, on the emulator's forward state
    cbs = []
    for c in cells:
        T, h = c["T"], c["h"]
        cbs.append(dict(T=T, B=B, h=h, gates=c["gates"], cs=c["cs"], W=c["W"], dh_all=g(T * B, h, seed=21 + h),
                        dh_last=g(B, h, seed=22 + h), dc_ext=g(T * B, h, seed=23 + h), dG=torch.zeros(T * B, 4 * h),
                        dc_scratch=torch.zeros(B, h)))
    dbs = [_to_dev(c) for c in cbs]
    EmuOps().lstm_bwd(cbs)
    ops.lstm_bwd(dbs)
    torch.cuda.synchronize()
    for c, d in zip(cbs, dbs):
        assert rel_l2(d["dG"], c["dG"]) < 1e-4, (c["h"], rel_l2(d["dG"], c["dG"]))


def test_lstm_variants_are_the_ones_intended(ops):
    """The per-variant launch counters of the ABI: which kernel served a cell is asserted, not assumed."""
    def run(h, B, force=0):
        c = _lstm_case(3, B, h, 3, seed=5)
        cb = dict(T=3, B=B, h=h, gates=c["gates"], cs=c["cs"], W=c["W"], dh_all=g(3 * B, h, seed=1), dh_last=None, dc_ext=None,
                  dG=torch.zeros(3 * B, 4 * h), dc_scratch=torch.zeros(B, h))
        EmuOps().lstm_fwd([c])
        EmuOps().lstm_bwd([cb])
        assert ops.lib.mfm_debug_lstm_force_nb(force) == 0
        assert ops.lib.mfm_debug_lstm_force_chains(2) == 0        # two chains per CTA: the shared-memory plan of the full step
        try:
            before = _variant_counts(ops)
            cg, cbg = _to_dev(c), _to_dev(cb)
            ops.lstm_fwd([cg])
            ops.lstm_bwd([cbg])
            torch.cuda.synchronize()
            after = _variant_counts(ops)
        finally:
            ops.lib.mfm_debug_lstm_force_nb(0)
            ops.lib.mfm_debug_lstm_force_chains(0)
        assert rel_l2(cg["hs"], c["hs"]) < 1e-4 and rel_l2(cg["gates"], c["gates"]) < 1e-4
        assert rel_l2(cbg["dG"], cb["dG"]) < 1e-4
        return [a - b for a, b in zip(after, before)]
    assert run(88, 300) == [1, 0, 1, 0, 0, 0]                 # wide (32-row) chains both ways
    assert run(88, 300, force=16) == [0, 1, 0, 1, 0, 0]       # half-width (16-row) chains forced
    assert run(104, 300) == [1, 0, 0, 1, 0, 0]                # backward of h = 104 only fits with 16-row chains
    assert run(48, 300, force=16) == [0, 1, 0, 1, 0, 0]       # two-copy layout: 64-row chains wide, 32-row half-width
    assert run(24, 300) == [1, 0, 1, 0, 0, 0] and run(24, 300, force=16) == [0, 1, 0, 1, 0, 0]
    assert run(200, 50) == [0, 0, 0, 0, 1, 1]                 # h > 128: CUDA-core kernels


def test_lstm_gx_leading_dimension(ops):
    """gx as a column block of a wider matrix (ld_gx): two cells fed by ONE input-projection GEMM."""
    T, B = 4, 37
    c1, c2 = _lstm_case(T, B, 32, T, seed=1), _lstm_case(T, B, 88, T, seed=2)
    wide = torch.cat([c1["gx"], c2["gx"]], 1).cuda()
    d1, d2 = _to_dev(c1), _to_dev(c2)
    d1["gx"], d2["gx"] = wide[:, :128], wide[:, 128:]
    EmuOps().lstm_fwd([c1, c2])
    ops.lstm_fwd([d1, d2])
    torch.cuda.synchronize()
    for c, d in ((c1, d1), (c2, d2)):
        assert rel_l2(d["hs"], c["hs"]) < 1e-4 and rel_l2(d["gates"], c["gates"]) < 1e-4


@pytest.mark.parametrize("T,B,h,dec,t0", [(6, 40, 88, False, 3), (5, 33, 24, True, 2), (4, 70, 104, True, 1), (7, 19, 130, False, 4)])
def test_lstm_bwd_split_in_time(ops, T, B, h, dec, t0):
    """A backward recurrence launched as two halves in time -- steps [t0, T) handing (dh, dc) to steps [0, t0) -- equals the
    single launch: tensor-core kernel (h <= 128) and CUDA-core kernel (h = 130), with dc_ext / dc_ext2 or dh_all."""
    H4 = 4 * h
    gates = torch.sigmoid(g(T * B, H4, seed=1)).cuda()
    cs = g((T + 1) * B, h, seed=2).cuda()
    cs[:B] = 0
    W = g(H4, h, seed=3, scale=0.2).cuda()
    base = dict(T=T, B=B, h=h, gates=gates, cs=cs, W=W, dh_all=g(T * B, h, seed=4).cuda() if dec else None,
                dh_last=None if dec else g(B, h, seed=5).cuda(), dc_ext=None if dec else g(T * B, h, seed=6).cuda(),
                dc_ext2=None if dec else g((T - 1) * B, h, seed=7).cuda(), dc_scratch=torch.zeros(B, h).cuda())
    one = dict(base, dG=torch.zeros(T * B, H4).cuda())
    ops.lstm_bwd([one])
    dG2 = torch.zeros(T * B, H4).cuda()
    dho, dco = torch.zeros(B, h).cuda(), torch.zeros(B, h).cuda()

    def sub(lo, hi):
        d = dict(base, T=hi - lo, gates=gates[lo * B:hi * B], cs=cs[lo * B:(hi + 1) * B], dG=dG2[lo * B:hi * B])
        for k in ("dh_all", "dc_ext"):
            if base[k] is not None:
                d[k] = base[k][lo * B:hi * B]
        if base["dc_ext2"] is not None:
            d["dc_ext2"] = base["dc_ext2"][lo * B:min(hi, T - 1) * B]
            d["dc_ext2_full"] = hi < T
        if hi == T:
            d.update(dh_out=dho, dc_out=dco)
        else:
            d.update(dh_last=dho, dc_last=dco)
        return d
    ops.lstm_bwd([sub(t0, T)])
    ops.lstm_bwd([sub(0, t0)])
    torch.cuda.synchronize()
    assert rel_l2(dG2, one["dG"]) < 2e-5
    # and against the torch statement
    ref = {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in base.items()}
    ref["dG"] = torch.zeros(T * B, H4)
    EmuOps().lstm_bwd([ref])
    assert rel_l2(dG2, ref["dG"]) < 1e-4


@pytest.mark.parametrize("simt", [False, True])
@pytest.mark.parametrize("T,B,mem,g1,g2,drop", [(3, 5, 9, 12, 13, False), (4, 21, 64, 128, 128, True), (2, 7, 300, 256, 32, False),
                                                (20, 64, 64, 128, 128, False), (1, 16, 32, 128, 40, True), (5, 161, 64, 100, 128, True),
                                                (3, 33, 16, 16, 16, False)])
def test_mfn_mem_fwd_bwd(ops, T, B, mem, g1, g2, drop, simt):
    """Both forms of the recurrence: tcgen05 (csrc/mem_ws.cu; mem <= 64, g <= 128) and CUDA cores (csrc/mfn.cu)."""
    ops.lib.mfm_debug_mem_force_simt(1 if simt else 0)
    n0 = [ops.lib.mfm_debug_mem_ws_count(i) for i in (0, 1)]
    try:
        _mfn_mem_case(ops, T, B, mem, g1, g2, drop)
    finally:
        ops.lib.mfm_debug_mem_force_simt(0)
    used = [ops.lib.mfm_debug_mem_ws_count(i) - n0[i] for i in (0, 1)]
    want = 1 if (not simt and mem <= 64 and g1 <= 128 and g2 <= 128) else 0
    assert used == [want, want], (used, want)


def _mfn_mem_case(ops, T, B, mem, g1, g2, drop):
    TB = T * B
    Wg1, Wg2 = g(g1, 20 + mem, seed=1, scale=0.2), g(g2, 20 + mem, seed=2, scale=0.2)
    rng = torch.tensor([5, 2], dtype=torch.int64)
    a = dict(T=T, B=B, mem=mem, g1=g1, g2=g2, G1pre=g(TB, g1, seed=3), G2pre=g(TB, g2, seed=4),
             cHat=torch.tanh(g(TB, mem, seed=5)), W1m=Wg1[:, 20:], W2m=Wg2[:, 20:],
             W12=g(mem, g1, seed=6, scale=0.2), b12=g(mem, seed=7), W22=g(mem, g2, seed=8, scale=0.2), b22=g(mem, seed=9),
             mems=torch.zeros((T + 1) * B, mem), U1=torch.zeros(TB, g1), U2=torch.zeros(TB, g2),
             Gam1=torch.zeros(TB, mem), Gam2=torch.zeros(TB, mem),
             drop1=(0.5, 3) if drop else None, drop2=(0.3, 4) if drop else None, rng=rng)
    ad = _to_dev(a)
    ad["W1m"], ad["W2m"] = Wg1.cuda()[:, 20:], Wg2.cuda()[:, 20:]
    EmuOps().mfn_mem_fwd(a)
    ops.mfn_mem_fwd(ad)
    torch.cuda.synchronize()
    for k in ("mems", "U1", "U2", "Gam1", "Gam2"):
        assert rel_l2(ad[k], a[k]) < TOL, k
    b = dict(a)
    b.update(scale1=2.0 if drop else 1.0, scale2=1.0 / 0.7 if drop else 1.0, dmem_last=g(B, mem, seed=10),
             dU1=torch.zeros(TB, g1), dU2=torch.zeros(TB, g2), dP1=torch.zeros(TB, mem), dP2=torch.zeros(TB, mem),
             dPc=torch.zeros(TB, mem))
    bd = _to_dev(b)
    bd["W1m"], bd["W2m"] = Wg1.cuda()[:, 20:], Wg2.cuda()[:, 20:]
    EmuOps().mfn_mem_bwd(b)
    ops.mfn_mem_bwd(bd)
    torch.cuda.synchronize()
    for k in ("dU1", "dU2", "dP1", "dP2", "dPc"):
        assert rel_l2(bd[k], b[k]) < 5e-5, k


@pytest.mark.parametrize("M,N", [(1, 1), (7, 30), (100, 400), (33, 1000)])
def test_softmax_gate(ops, M, N):
    L, cs = g(M, N, seed=1, scale=3.0), g(M, N, seed=2)
    att_cpu, att_gpu = L.clone(), L.clone().cuda()
    o_cpu, o_gpu = torch.zeros(M, N), torch.zeros(M, N).cuda()
    EmuOps().softmax_gate_fwd(att_cpu, cs, o_cpu)
    ops.softmax_gate_fwd(att_gpu, cs.cuda(), o_gpu)
    assert rel_l2(att_gpu, att_cpu) < TOL and rel_l2(o_gpu, o_cpu) < TOL
    dA = g(M, N, seed=3)
    dL_c, dc_c = torch.zeros(M, N), torch.zeros(M, N)
    dL_g, dc_g = torch.zeros(M, N).cuda(), torch.zeros(M, N).cuda()
    EmuOps().softmax_gate_bwd(dA, att_cpu, cs, dL_c, dc_c)
    ops.softmax_gate_bwd(dA.cuda(), att_cpu.cuda(), cs.cuda(), dL_g, dc_g)
    assert float((dL_g.cpu() - dL_c).abs().max()) < 5e-5 * max(1e-3, float(dL_c.abs().max())) + 1e-7
    assert rel_l2(dc_g, dc_c) < TOL


@pytest.mark.parametrize("B,dim", [(1, 1), (6, 3), (33, 8), (100, 80), (257, 32), (64, 256)])
def test_mmd(ops, B, dim):
    zfull = g(B, dim + 5, seed=1)
    z, n = zfull[:, 2:2 + dim], g(B, dim, seed=2)
    o_cpu, o_gpu = torch.zeros(1), torch.full((1,), 5.0).cuda()
    EmuOps().mmd_fwd(z, n, o_cpu)
    ops.mmd_fwd(zfull.cuda()[:, 2:2 + dim], n.cuda(), o_gpu)
    assert abs(float(o_gpu) - float(o_cpu)) < 1e-5 * max(1.0, abs(float(o_cpu))), (float(o_gpu), float(o_cpu))
    dz_c, dz_g = torch.ones(B, dim), torch.ones(B, dim).cuda()
    EmuOps().mmd_bwd(z, n, 0.7, dz_c)
    ops.mmd_bwd(zfull.cuda()[:, 2:2 + dim], n.cuda(), 0.7, dz_g)
    assert rel_l2(dz_g - 1.0, dz_c - 1.0) < 1e-4
    sd = torch.tensor([0.5]).cuda()
    dz_g2 = torch.ones(B, dim).cuda()
    ops.mmd_bwd(zfull.cuda()[:, 2:2 + dim], n.cuda(), 1.4, dz_g2, scale_dev=sd)
    assert rel_l2(dz_g2, dz_g) < 1e-6


def test_mmd_gemm_formulation(ops):
    """rownorm2 + GEMM + mmd_kexp (+ colsum, GEMMs, mmd_combine) against the direct pairwise statement."""
    emu = EmuOps()
    B, dim = 300, 80
    z, n = g(B, dim, seed=1), g(B, dim, seed=2)
    ref, dz_ref = torch.zeros(1), torch.zeros(B, dim)
    emu.mmd_fwd(z, n, ref)
    emu.mmd_bwd(z, n, 0.7, dz_ref)
    zd, nd = z.cuda(), n.cuda()
    nz, ng = torch.zeros(B).cuda(), torch.zeros(B).cuda()
    ops.rownorm2(zd, nz)
    ops.rownorm2(nd, ng)
    assert rel_l2(nz, (z * z).sum(1)) < 1e-6
    slot = torch.zeros(1).cuda()
    Kzz, Kgz, Kgg = (torch.zeros(B, B).cuda() for _ in range(3))
    ib = 1.0 / (B * B)
    ops.gemm("nt", zd, zd, Kzz); ops.mmd_kexp(Kzz, nz, nz, dim, ib, slot)
    ops.gemm("nt", nd, zd, Kgz); ops.mmd_kexp(Kgz, ng, nz, dim, -2 * ib, slot)
    ops.gemm("nt", nd, nd, Kgg); ops.mmd_kexp(Kgg, ng, ng, dim, ib, slot)
    assert abs(float(slot) - float(ref)) < 2e-5 * max(1.0, abs(float(ref))), (float(slot), float(ref))
    rc = torch.zeros(2 * B).cuda()
    ops.colsum(Kzz, rc[:B]); ops.colsum(Kgz, rc[B:])
    t1, t2 = torch.zeros(B, dim).cuda(), torch.zeros(B, dim).cuda()
    ops.gemm("nn", Kzz, zd, t1); ops.gemm("tn", Kgz, nd, t2)
    dz = torch.zeros(B, dim).cuda()
    ops.mmd_combine(zd, rc[:B], rc[B:], t1, t2, 0.7, dz)
    assert rel_l2(dz, dz_ref) < 2e-4, rel_l2(dz, dz_ref)


def test_small_kernels(ops):
    emu = EmuOps()
    src = g(37, 50, seed=1)
    for acc in (False, True):
        d_c, d_g = torch.ones(37, 64), torch.ones(37, 64).cuda()
        emu.copy2d(src[:, 5:25], d_c[:, 10:30], accumulate=acc)
        ops.copy2d(src.cuda()[:, 5:25], d_g[:, 10:30], accumulate=acc)
        assert torch.equal(d_g.cpu(), d_c)
    a, b = g(1000, seed=2), g(1000, seed=3)
    o = torch.zeros(1000).cuda()
    ops.add(a.cuda(), b.cuda(), o)
    assert torch.equal(o.cpu(), a + b)
    ops.zero(o)
    assert float(o.abs().max()) == 0.0
    A = g(5000, 77, seed=4)
    out_c, out_g = torch.ones(77), torch.ones(77).cuda()
    emu.colsum(A[:, 3:70], out_c[:67])
    ops.colsum(A.cuda()[:, 3:70], out_g[:67])
    assert rel_l2(out_g, out_c) < 1e-5
    dy, y = g(33, 21, seed=5), torch.relu(g(33, 21, seed=6))
    r_c, r_g = torch.zeros(33, 21), torch.zeros(33, 21).cuda()
    emu.relu_bwd(dy, y, r_c)
    ops.relu_bwd(dy.cuda(), y.cuda(), r_g)
    assert torch.equal(r_g.cpu(), r_c)


def test_loss_heads_and_adam(ops):
    emu = EmuOps()
    xh, x = g(640, 20, seed=1), g(640, 45, seed=2)
    s_c, s_g = torch.zeros(1), torch.zeros(1).cuda()
    d_c, d_g = torch.zeros(640, 20), torch.zeros(640, 20).cuda()
    emu.mse_fwd_bwd(xh, x[:, 5:25], 1.0 / 12800, 2.0 * 0.5 / 12800, s_c, d_c)
    ops.mse_fwd_bwd(xh.cuda(), x.cuda()[:, 5:25], 1.0 / 12800, 2.0 * 0.5 / 12800, s_g, d_g)
    assert abs(float(s_g) - float(s_c)) < 1e-5 * float(s_c) and rel_l2(d_g, d_c) < 1e-6
    yh, y = g(77, 4, seed=3), g(77, 4, seed=4)
    s_c, s_g = torch.zeros(1), torch.zeros(1).cuda()
    d_c, d_g = torch.zeros(77, 4), torch.zeros(77, 4).cuda()
    emu.l1_fwd_bwd(yh, y, 1.0 / 308, s_c, d_c)
    ops.l1_fwd_bwd(yh.cuda(), y.cuda(), 1.0 / 308, s_g, d_g)
    assert abs(float(s_g) - float(s_c)) < 1e-5 * float(s_c) and rel_l2(d_g, d_c) < 1e-6
    lab = torch.randint(0, 4, (77,), generator=torch.Generator().manual_seed(1))
    s_c, s_g = torch.zeros(1), torch.zeros(1).cuda()
    emu.ce_fwd_bwd(yh, lab, 1.0 / 77, s_c, d_c)
    ops.ce_fwd_bwd(yh.cuda(), lab.cuda(), 1.0 / 77, s_g, d_g)
    assert abs(float(s_g) - float(s_c)) < 1e-5 * float(s_c) and rel_l2(d_g, d_c) < 1e-5
    lb_c = torch.arange(16, dtype=torch.float32) * 0.1
    lb_g = lb_c.clone().cuda()
    emu.loss_total(lb_c, 1.0, 0.01, 0.5, 0.7)
    ops.loss_total(lb_g, 1.0, 0.01, 0.5, 0.7)
    assert abs(float(lb_g[8]) - float(lb_c[8])) < 1e-6
    # Adam: three steps against torch.optim.Adam itself
    p0, grads = g(5000, seed=5), [g(5000, seed=6 + i) for i in range(3)]
    pt = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([pt])
    pg, m, v = p0.clone().cuda(), torch.zeros(5000).cuda(), torch.zeros(5000).cuda()
    st = torch.tensor([1e-3, 0, 0, 0], dtype=torch.float32).cuda()
    for gi in grads:
        pt.grad = gi.clone()
        opt.step()
        ops.adam(pg, (gi * 4.0).cuda(), m, v, st, grad_scale=0.25)
    assert rel_l2(pg - p0.cuda(), pt.detach() - p0) < 1e-4
    assert float(st[1]) == 3.0


def test_randn_and_rng(ops):
    rng = torch.tensor([42, 0], dtype=torch.int64)
    rg = rng.clone().cuda()
    ops.rng_tick(rg)
    assert rg.cpu().tolist() == [42, 1]
    out_g, out_c = torch.zeros(100000).cuda(), torch.zeros(100000)
    ops.randn(out_g, rg, 20)
    EmuOps().randn(out_c, rg.cpu(), 20)
    assert float((out_g.cpu() - out_c).abs().max()) < 1e-4
    assert abs(float(out_g.mean())) < 0.02 and abs(float(out_g.std()) - 1.0) < 0.02
    k = keep_mask(rg.cpu(), 7, 0.3, 100, 50)
    assert abs(float(k.mean()) - 0.7) < 0.03
    # consecutive steps draw unrelated samples (the first RNG produced the step-s sample shifted by one element at step s+2)
    ops.rng_tick(rg)
    ops.rng_tick(rg)
    out2 = torch.zeros(100000).cuda()
    ops.randn(out2, rg, 20)
    a, b = out_g.cpu(), out2.cpu()
    for shift in (0, 1, 2):
        cc = float(torch.corrcoef(torch.stack([a[shift:50000 + shift], b[:50000]]))[0, 1])
        assert abs(cc) < 0.02, (shift, cc)


def test_bad_arguments_fail_loudly(ops):
    from factorized_b200.cuda_ops import MfmCudaError
    with pytest.raises(MfmCudaError):
        ops.gemm("nt", torch.zeros(4, 4), torch.zeros(4, 4).cuda(), torch.zeros(4, 4).cuda())     # CPU tensor
    with pytest.raises(MfmCudaError):
        ops.gemm("nt", torch.zeros(4, 5).cuda(), torch.zeros(4, 4).cuda(), torch.zeros(4, 4).cuda())  # K mismatch
    with pytest.raises(MfmCudaError):
        ops.mmd_fwd(torch.zeros(4, 300).cuda(), torch.zeros(4, 300).cuda(), torch.zeros(1).cuda())   # dim > 256
