"""Hyper-parameter sets and benchmark workloads of the MFM path.

``best_acc_configs`` is the only fixed MFM hyper-parameter set in the reference (mfm_mosi.py:1239-1286, function
``best_acc``); the workloads are BASELINE.json's ``configs`` (SURVEY.md section 8d spells out the shapes; the IEMOCAP and
POM feature dimensions are the survey's assumptions -- the reference has no script for them).
"""


def best_acc_configs(input_dims=(300, 5, 20), output_dim=1, dropout=True):
    dr = (lambda p: p) if dropout else (lambda p: 0.0)
    config = dict(
        input_dims=list(input_dims), h_dims=[88, 64, 48],
        zy_size=32, zl_size=32, za_size=8, zv_size=80,
        fy_size=16, fl_size=88, fa_size=8, fv_size=8, memsize=64,
        zy_to_fy_dropout=dr(0.0), zl_to_fl_dropout=dr(0.2), za_to_fa_dropout=dr(0.2),
        zv_to_fv_dropout=dr(0.7), fy_to_y_dropout=dr(0.0),
        lda_mmd=1.0, lda_xl=1.0, lda_xa=0.01, lda_xv=0.5,
        missing=0, windowsize=2, batchsize=32, num_epochs=30, lr=0.01, momentum=0.9,
        output_dim=output_dim, type="mfm",
    )
    nn_ = lambda s: dict(shapes=s, drop=dr(0.5))
    return [config, nn_(128), nn_(128), nn_(128), nn_(128), nn_(64)]


# name -> (input_dims, T, head, output_dim, default batch, batch is "per_gpu" | "global", description)
WORKLOADS = {
    "mosi": dict(input_dims=(300, 5, 20), T=20, head="l1", out=1, batch=2048, batch_is="per_gpu",
                 desc="synthetic CMU-MOSI shapes: text 300 / audio 5 / visual 20, T=20 (BASELINE configs[1], headline batch 2048)"),
    "mosei": dict(input_dims=(300, 74, 35), T=50, head="l1", out=1, batch=512, batch_is="global",
                  desc="synthetic CMU-MOSEI shapes: text 300 / audio 74 / visual 35, T=50, global batch 512 (BASELINE configs[2])"),
    "iemocap": dict(input_dims=(300, 74, 35), T=20, head="ce", out=4, batch=256, batch_is="per_gpu",
                    desc="IEMOCAP-like shapes 300/74/35, T=20, 4-class cross-entropy head, batch 256 (BASELINE configs[3])"),
    "pom": dict(input_dims=(300, 43, 43), T=100, head="l1", out=16, batch=1024, batch_is="per_gpu",
                desc="POM-like shapes 300/43/43, T=100, 16 regression targets, batch 1024 per GPU (BASELINE configs[4])"),
}
