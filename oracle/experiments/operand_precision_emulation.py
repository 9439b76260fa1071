"""TEST INFRASTRUCTURE / EVIDENCE -- what operand precision the hGRU's 15x15 convolutions need.

Emulates on the CPU (torch, float64 accumulation) the state error of the 8-step recurrence at BASELINE configs[1]
(25 channels, 64x64, T = 8, stress weights) when ONLY the operands of the two horizontal convolutions are rounded,
exactly as a tensor-core path would round them: bf16 (8 significand bits), tf32 (11), fp16 (11), bf16 hi+lo (16).
Everything else (gates, state, accumulation) stays float64, so the numbers are the floor a kernel of that operand
type can reach.  Answers BASELINE.json configs[1] "fp32 vs tf32": a tf32 arm cannot meet the 1e-4 bar of the fp32
path, bf16x3 can.  Run: python oracle/experiments/operand_precision_emulation.py  (writes nothing; prints a table)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from monkey_pose_b200 import initialization as init  # noqa: E402
from oracle import hgru_oracle_torch as otorch  # noqa: E402


def round_bits(x, bits):
    """Round-to-nearest-even to `bits` significand bits (incl. the implicit one), float64 in / out."""
    m, e = torch.frexp(x)
    return torch.ldexp(torch.round(m * (1 << bits)) / (1 << bits), e)


def rounders():
    return {
        "bf16 (8 bits)": lambda t: t.to(torch.bfloat16).to(torch.float64),
        "tf32 (11 bits)": lambda t: round_bits(t, 11),
        "fp16 (11 bits)": lambda t: t.to(torch.float16).to(torch.float64),
        "bf16 hi+lo (16 bits)": lambda t: (lambda hi: hi + (t - hi).to(torch.bfloat16).to(torch.float64))(
            t.to(torch.bfloat16).to(torch.float64)),
    }


def forward(X, H2, p, T, rnd):
    w = rnd(p["p_r"]) if rnd else p["p_r"]
    states = []
    for t in range(T):
        G1 = torch.sigmoid(otorch._conv_same(H2, p["i_r"]) + p["i_b"])
        A = H2 * G1
        C1 = otorch._conv_same(rnd(A) if rnd else A, w) + p["lateral_bias"]
        H1 = torch.tanh(X - (p["beta"] * H2 + p["nu"]) * C1)
        G2 = torch.sigmoid(otorch._conv_same(H1, p["o_r"]) + p["o_b"])
        C2 = otorch._conv_same(rnd(H1) if rnd else H1, w) + p["lateral_bias"]
        e = p["gamma"] * C2
        Ht = torch.tanh(p["kappa"] * (H1 + e) + p["omega"] * (H1 * e))
        H2 = (G2 * H2 + (1.0 - G2) * Ht) * p["rho"][t]
        states.append((H1, H2))
    return states


def main():
    k, T = 25, 8
    rng = np.random.default_rng(17 + k)
    X = torch.as_tensor(rng.uniform(-1, 1, size=(2, 64, 64, k))).permute(0, 3, 1, 2).contiguous()
    O0 = torch.as_tensor(init.hidden_init((2, 64, 64, k), seed=3, limit=0.5)).double().permute(0, 3, 1, 2).contiguous()
    p = otorch._prep_hgru(init.hgru_params(k, 15, T, seed=9, stress=5.0), torch.float64)
    with torch.no_grad():
        ref = forward(X, O0, p, T, None)
        print("%-24s %14s %14s   (max|a-b| / max|b|, worst over H1_t, H2_t, t = 0..%d)" % ("conv operands", "worst", "final H2", T - 1))
        for name, rnd in rounders().items():
            got = forward(X, O0, p, T, rnd)
            worst = 0.0
            for (a1, a2), (b1, b2) in zip(got, ref):
                worst = max(worst, float((a1 - b1).abs().max() / b1.abs().max()), float((a2 - b2).abs().max() / b2.abs().max()))
            fin = float((got[-1][1] - ref[-1][1]).abs().max() / ref[-1][1].abs().max())
            print("%-24s %14.3e %14.3e   %s" % (name, worst, fin, "meets 1e-4" if worst < 1e-4 else ("meets 1e-2" if worst < 1e-2 else "fails")))


if __name__ == "__main__":
    main()
