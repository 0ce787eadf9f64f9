"""DESIGN.md section 4's precision policy for the text side, pinned on CPU: the three-term bf16 product keeps logw within 5e-5 of
fp32 (no duration changes), plain bf16 operands do not (tests/probe_text_precision.py; full run: profiles/r02_text_precision_probe.md)."""
from probe_text_precision import run


def test_bf16x3_is_fp32_faithful_and_plain_bf16_is_not():
    ids, rows = run(n_utts=6, seed=1)
    r = {m: (e, f) for m, e, f in rows}
    assert ids > 500
    assert r["bf16x3"][0] < 5e-5 and r["bf16x3"][1] == 0
    assert r["bf16"][0] > 1e-3                    # an order of magnitude past what ceil(exp(logw)) tolerates
    assert r["tf32"][0] > 5 * r["bf16x3"][0]
