"""Host restatement of the product's counter-based RNG (factorized_b200/csrc/common.cuh: site_key, rng_bits, drop_keep)
-- TEST INFRASTRUCTURE, like the rest of oracle/.

The reference draws its dropout masks from torch's generator, which no other implementation can reproduce; parity of a
train-mode step is therefore checked by REPLAYING the masks the CUDA step drew (a pure function of seed, step, site and
element index) inside the oracle (mfm_oracle.dropout(mask=...)).  Used by tests/ and by bench.py's parity check.
"""
import torch

M32 = 0xFFFFFFFF


def _fmix32(h):
    h = h & M32
    h = h ^ (h >> 16)
    h = (h * 0x85EBCA6B) & M32
    h = h ^ (h >> 13)
    h = (h * 0xC2B2AE35) & M32
    h = h ^ (h >> 16)
    return h


def site_seed(rng, site):
    """common.cuh::site_key -- every input passes its own mixing round (consecutive steps give unrelated keys)."""
    seed, step = int(rng[0]) & M32, int(rng[1]) & M32
    return _fmix32(_fmix32(seed ^ _fmix32((step + 0x9E3779B9) & M32)) + ((site * 0x7F4A7C15) & M32))


def rng_bits(sseed, idx):
    """common.cuh::rng_bits -- counter -> 32 bits, two rounds, the key enters both (idx: int64 tensor)."""
    return _fmix32(_fmix32(((idx & M32) + sseed) & M32) ^ sseed)


def keep_mask(rng, site, p, rows, cols, row0=0):
    """keep[m,n] = u(idx) >= p with idx = (row0+m)*cols + n (32-bit wrap)."""
    idx = (torch.arange(rows, dtype=torch.int64).view(-1, 1) + row0) * cols + torch.arange(cols, dtype=torch.int64).view(1, -1)
    h = rng_bits(site_seed(rng, site), idx)
    u = (h >> 8).to(torch.float32) * (1.0 / 16777216.0)
    return (u >= p).to(torch.float32)


# dropout sites of the training schedule (factorized_b200/engine.py: SITE_*)
SITES = dict(att1=1, att2=2, gamma1=3, gamma2=4, fy=5, fl=6, fa=7, fv=8, y=9)


def _missing_masks_and_branches(eng, rng_cpu, site):
    """MFM_missing: the MFN sites as in MFM; the generative half's sites once per pass p (site index + 32 p, stashes suffixed
    "@p"), keyed "<site>@p" for oracle.mfm_missing_forward (pass 0 unsuffixed)."""
    ws, dm = eng.ws, eng.dm
    T, n = dm.T, dm.B
    site("att1", "att1", dm.p_att1, ws["H1"], T * n, (T, n, -1))
    site("att2", "att2", dm.p_att2, ws["H2"], T * n, (T, n, -1))
    site("gamma1", "gamma1", dm.p_g1, ws["U1"], T * n, (T, n, -1))
    site("gamma2", "gamma2", dm.p_g2, ws["U2"], T * n, (T, n, -1))
    masks, br = site.masks, site.br
    for p in range(4):
        sfx = "" if p == 0 else "@%d" % p
        site("fy", "fy" + sfx + "1", dm.p_fy, ws["F1y@%d" % p], n, None, 32 * p, "fy" + sfx)
        site("y", "y1" + sfx, dm.p_y, ws["Y1@%d" % p], n, None, 32 * p, "y" + sfx)
        br["fy" + sfx] = (ws["FY@%d" % p] > 0).float().cpu()
        for m, tag in enumerate("lav"):
            site("f" + tag, "f%s%s1" % (tag, sfx), dm.p_f[m], ws["F1_%d@%d" % (m, p)], n, None, 32 * p, "f" + tag + sfx)
            br["f" + tag + sfx] = (ws["EMB%d@%d" % (m, p)][:, dm.fy:] > 0).float().cpu()
    return masks, br


def train_masks_and_branches(eng, rng_cpu):
    """Dropout keep-masks of a CUDA train-mode step, regenerated on the CPU from the step's RNG state, and the ReLU
    branches read back from the CUDA stashes: a kept unit's branch is (output > 0); a dropped unit's branch was not
    observed (-1: the oracle decides itself).  `eng` is the step's factorized_b200.engine.Engine (its workspace)."""
    ws, dm = eng.ws, eng.dm
    T, n = dm.T, dm.B
    masks, br = {}, {}

    def site(key, bkey, p, buf, rows, shape3=None, site_off=0, mkey=None):
        out = buf.detach().cpu()
        taken = (out > 0).float()
        if p > 0.0:
            k = keep_mask(rng_cpu, SITES[key] + site_off, p, rows, out.shape[1])
            taken = torch.where(k > 0, taken, torch.full_like(taken, -1.0))
            masks[mkey or key] = k.view(shape3) if shape3 else k
        br[bkey] = taken.view(shape3) if shape3 else taken
    site.masks, site.br = masks, br
    if hasattr(eng, "cross"):                               # factorized_b200.missing.MissingEngine: four generative passes
        return _missing_masks_and_branches(eng, rng_cpu, site)
    abl = getattr(eng, "abl", None)                         # factorized_b200.ablations.AblationEngine: M_A .. M_D
    has_mfn = eng.has_mfn if abl else not getattr(eng, "ef", False)
    if has_mfn:                                             # (MFM_KL_EF, M_B and M_D have no MFN: no attention / gamma dropouts)
        site("att1", "att1", dm.p_att1, ws["H1"], T * n, (T, n, -1))
        site("att2", "att2", dm.p_att2, ws["H2"], T * n, (T, n, -1))
        site("gamma1", "gamma1", dm.p_g1, ws["U1"], T * n, (T, n, -1))
        site("gamma2", "gamma2", dm.p_g2, ws["U2"], T * n, (T, n, -1))
    if abl is None or has_mfn:
        site("fy", "fy1", dm.p_fy, ws["F1y"], n)
        br["fy"] = (ws["FY"] > 0).float().cpu()
    if abl != "m_d":
        site("y", "y1", dm.p_y, ws["Y1"], n)
    for m, tag in enumerate("lav"):
        if abl is None:
            site("f" + tag, "f%s1" % tag, dm.p_f[m], ws["F1_%d" % m], n)
            br["f" + tag] = (ws["EMB%d" % m][:, dm.fy:] > 0).float().cpu()
        elif m in eng.fdst:
            site("f" + tag, "f%s1" % tag, dm.p_f[m], ws["F1_%d" % m], n)
            br["f" + tag] = (eng.fdst[m] > 0).float().cpu()
    return masks, br
