#!/usr/bin/env python
"""Design study (CPU, float32 emulation): rounding error of the F(2,2) fast-FIR split of the polyphase DDC against the
float64 windowed oracle, next to the direct float32 form the kernels used before.  Run: python tools/winograd_error.py

    y[2r]   = M1 + M2,   y[2r+1] = M2 - M3
    M1 = sum_i (x_b[2(r+i)]   - x_b[2(r+i)+1]) c[2i]          (block index b, per phase)
    M2 = sum_i  x_b[2(r+i)+1] (c[2i] + c[2i+1])
    M3 = sum_i (x_b[2(r+i)+1] - x_b[2(r+i)+2]) c[2i+1]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dc_sand_b200 import synth, taps  # noqa: E402
from oracle import ddc_oracle as orc  # noqa: E402


def folded(tp, step, D):
    T = len(tp)
    J = -(-T // D)
    J += J & 1
    k = np.arange(J * D)
    h = np.zeros(J * D)
    h[:T] = tp[::-1] / tp.sum()
    return (h * np.exp(-2j * np.pi * ((k * step) % 1.0))), J


def run(T=256, D=16, n=1 << 18, fc=100e6, fs=1712e6, float_input=False):
    tp = taps.coefficients("ddc_coeff_107MHz.csv") if T == 256 else __import__("scipy.signal").signal.firwin(T, 0.8 / D)
    x = synth.digitiser_stream(n, 1234).astype(np.float32)
    if float_input:
        x = (x * np.float32(0.3713)).astype(np.float32)
    step = orc.phase_step_cycles(n, fc, fs)
    c, J = folded(tp, step, D)
    M = (n - T) // D + 1
    M -= M & 1
    M = min(M, 8192)
    ref = orc.ddc_windowed_f64(x, 0, M, step, tp, D)
    rot = np.exp(-2j * np.pi * ((np.arange(M) * D * step) % 1.0))
    c32 = c.astype(np.complex64)
    # direct float32: accumulate tap by tap in float32 (phase-major order as in the kernel)
    xb = np.zeros((M + J + 2) * D, np.float32)
    xb[: min(len(x), len(xb))] = x[: len(xb)]
    xb = xb.reshape(-1, D)
    acc = np.zeros((M, 2), np.float32)
    for d in range(D):
        for j in range(J):
            t = c32[j * D + d]
            s = xb[j:j + M, d]
            acc[:, 0] += s * np.float32(t.real)
            acc[:, 1] += s * np.float32(t.imag)
    y_direct = (acc[:, 0].astype(np.float64) + 1j * acc[:, 1]) * rot
    # F(2,2)
    ge = c[0::2 * D * 1].copy()  # placeholder (overwritten below)
    cm = c.reshape(J, D)
    ge = cm[0::2].astype(np.complex64)
    go = cm[1::2].astype(np.complex64)
    gs = (cm[0::2] + cm[1::2]).astype(np.complex64)
    H = M // 2
    m1 = np.zeros((H, 2), np.float32)
    m2 = np.zeros((H, 2), np.float32)
    m3 = np.zeros((H, 2), np.float32)
    for d in range(D):
        col = xb[:, d]
        de = col[0::2][: H + J // 2] - col[1::2][: H + J // 2]
        wo = col[1::2][: H + J // 2]
        do = col[1::2][: H + J // 2] - col[2::2][: H + J // 2]
        for i in range(J // 2):
            for acc_, seq, g in ((m1, de, ge), (m2, wo, gs), (m3, do, go)):
                t = g[i, d]
                s = seq[i:i + H]
                acc_[:, 0] += s * np.float32(t.real)
                acc_[:, 1] += s * np.float32(t.imag)
    y0 = (m1 + m2).astype(np.float32)
    y1 = (m2 - m3).astype(np.float32)
    yw = np.empty(M, np.complex128)
    yw[0::2] = y0[:, 0].astype(np.float64) + 1j * y0[:, 1]
    yw[1::2] = y1[:, 0].astype(np.float64) + 1j * y1[:, 1]
    yw *= rot
    s = np.abs(ref).max()
    for name, y in (("direct", y_direct), ("F(2,2)", yw)):
        print(f"T={T} D={D} float_input={float_input} {name:7s} max_err/max|ref| = {np.abs(y - ref).max() / s:.3e}   "
              f"rel_l2 = {np.linalg.norm(y - ref) / np.linalg.norm(ref):.3e}   (|M1|max/|y|max = {np.abs(m1).max() / s:.2f})")


if __name__ == "__main__":
    run()
    run(float_input=True)
    run(T=1024, D=16)
    run(T=512, D=32)


def run_nested(T=256, D=16, n=1 << 18, fc=100e6, fs=1712e6, float_input=False):
    """Two nested F(2,2) levels (9/16 of the multiplies): float32 emulation against the float64 oracle."""
    tp = taps.coefficients("ddc_coeff_107MHz.csv") if T == 256 else __import__("scipy.signal").signal.firwin(T, 0.8 / D)
    x = synth.digitiser_stream(n, 1234).astype(np.float32)
    if float_input:
        x = (x * np.float32(0.3713)).astype(np.float32)
    step = orc.phase_step_cycles(n, fc, fs)
    c, J = folded(tp, step, D)
    J4 = -(-J // 4) * 4
    cm = np.zeros((J4, D), np.complex128)
    cm[:J] = c.reshape(J, D)
    M = min(((n - T) // D + 1) // 4 * 4, 8192)
    ref = orc.ddc_windowed_f64(x, 0, M, step, tp, D)
    rot = np.exp(-2j * np.pi * ((np.arange(M) * D * step) % 1.0))
    xb = np.zeros((M + J4 + 4) * D, np.float32)
    xb[: min(len(x), len(xb))] = x[: len(xb)]
    xb = xb.reshape(-1, D)
    Q = M // 4
    f32 = np.float32

    def split_taps(g):      # g[i] -> (g[2i], g[2i]+g[2i+1], g[2i+1])
        return g[0::2], g[0::2] + g[1::2], g[1::2]

    def split_data(s):      # s[n] -> (s[2n]-s[2n+1], s[2n+1], s[2n+1]-s[2n+2]) in float32
        L = (len(s) - 1) // 2
        return (s[0:2 * L:2] - s[1:2 * L:2]).astype(f32), s[1:2 * L:2].astype(f32), (s[1:2 * L:2] - s[2:2 * L + 1:2]).astype(f32)

    acc = np.zeros((3, 3, Q, 2), f32)
    for d in range(D):
        col = xb[:, d]
        for a, (s1, g1) in enumerate(zip(split_data(col), split_taps(cm[:, d]))):
            for b, (s2, g2) in enumerate(zip(split_data(s1), split_taps(g1))):
                g2 = g2.astype(np.complex64)
                for i in range(J4 // 4):
                    seg = s2[i:i + Q]
                    acc[a, b, :, 0] += seg * f32(g2[i].real)
                    acc[a, b, :, 1] += seg * f32(g2[i].imag)
    # level-2 combine: M_a[2p] = A + B, M_a[2p+1] = B - C
    Ml = np.zeros((3, 2 * Q, 2), f32)
    for a in range(3):
        Ml[a, 0::2] = acc[a, 0] + acc[a, 1]
        Ml[a, 1::2] = acc[a, 1] - acc[a, 2]
    y = np.zeros((M, 2), f32)
    y[0::2] = Ml[0] + Ml[1]
    y[1::2] = Ml[1] - Ml[2]
    yw = (y[:, 0].astype(np.float64) + 1j * y[:, 1]) * rot
    s = np.abs(ref).max()
    print(f"T={T} D={D} float_input={float_input} nested  max_err/max|ref| = {np.abs(yw - ref).max() / s:.3e}   "
          f"rel_l2 = {np.linalg.norm(yw - ref) / np.linalg.norm(ref):.3e}")


if __name__ == "__main__":
    run_nested()
    run_nested(float_input=True)
    run_nested(T=1024, D=16)
