"""numpy emulation of the kernel-side semantics (channel-last implicit-GEMM conv with explicit
tap offsets, folded flips, interleaved gate, polyphase transposed conv) over the PACKED blobs.
Lets the packing logic be verified against the oracle on CPU, with no GPU."""
import numpy as np


def conv_cl(x, w, b, toff, n):
    """x: [L, Cin]; w: [taps, Cin, N4]; out[t, n] = sum_tap x[t + toff] @ w[tap] + b (zero outside)."""
    L = x.shape[0]
    out = np.zeros((L, w.shape[2]), np.float64)
    for k, off in enumerate(toff):
        xs = np.zeros_like(x, dtype=np.float64)
        lo, hi = max(0, -off), min(L, L - off)
        if hi > lo:
            xs[lo:hi] = x[lo + off:hi + off]
        out += xs @ w[k].astype(np.float64)
    if b is not None:
        out += b
    return out[:, :n].astype(np.float32)


def sym_taps(k, d):
    return [(i - (k - 1) // 2) * d for i in range(k)]


def lrelu(x, s):
    return np.where(x >= 0, x, x * s).astype(np.float32)


def flow_reverse(P, blobs, a, sid=None):
    H, C = a.hidden, a.inter
    half = C // 2
    P = P.copy()
    for s in range(len(a.flow_layers)):
        flipped = s % 2 == 0
        xcol, ocol = (half, 0) if flipped else (0, half)
        h = conv_cl(P[:, xcol:xcol + half], blobs[f"flow.{s}.pre.w"], blobs[f"flow.{s}.pre.b"], [0], H)
        skip = np.zeros_like(h)
        for i in range(a.wn_layers):
            d = a.wn_dilation_rate ** i
            xin = conv_cl(h, blobs[f"flow.{s}.in.{i}.w"], blobs[f"flow.{s}.in.{i}.b"], sym_taps(a.wn_kernel, d), 2 * H)
            if sid is not None:
                xin = xin + blobs[f"flow.{s}.cond_tab.{i}"][sid]
            acts = np.tanh(xin[:, 0::2]) * (1.0 / (1.0 + np.exp(-xin[:, 1::2])))
            nrs = 2 * H if i < a.wn_layers - 1 else H
            rs = conv_cl(acts.astype(np.float32), blobs[f"flow.{s}.rs.{i}.w"], blobs[f"flow.{s}.rs.{i}.b"], [0], nrs)
            if i < a.wn_layers - 1:
                h = h + rs[:, :H]
                skip = skip + rs[:, H:]
            else:
                skip = skip + rs
        m = conv_cl(skip, blobs[f"flow.{s}.post.w"], blobs[f"flow.{s}.post.b"], [0], half)
        P[:, ocol:ocol + half] = P[:, ocol:ocol + half] - m
    return P


def decoder(z, blobs, a, sid=None):
    C0 = a.up_init
    x = conv_cl(z, blobs["dec.pre.w"], blobs["dec.pre.b"], sym_taps(7, 1), C0)
    if sid is not None:
        x = x + blobs["dec.cond_tab"][sid]
    ch = C0
    nk = len(a.rb_kernels)
    for i, u in enumerate(a.up_rates):
        co = ch // 2
        xin = lrelu(x, 0.1)
        L = xin.shape[0]
        X = np.zeros((L, u * co), np.float32)
        X[:, : (u // 2) * co] = conv_cl(xin, blobs[f"dec.ups.{i}.A.w"], blobs[f"dec.ups.{i}.A.b"], [-1, 0], (u // 2) * co)
        X[:, (u // 2) * co:] = conv_cl(xin, blobs[f"dec.ups.{i}.B.w"], blobs[f"dec.ups.{i}.B.b"], [0, 1], (u // 2) * co)
        X = X.reshape(L * u, co)
        xs = None
        for j in range(nk):
            n = i * nk + j
            cur = X
            for c, d in enumerate(a.rb_dilations[j]):
                k = a.rb_kernels[j]
                if a.resblock == "1":
                    t1 = conv_cl(lrelu(cur, 0.1), blobs[f"dec.rb.{n}.c1.{c}.w"], blobs[f"dec.rb.{n}.c1.{c}.b"], sym_taps(k, d), co)
                    cur = conv_cl(lrelu(t1, 0.1), blobs[f"dec.rb.{n}.c2.{c}.w"], blobs[f"dec.rb.{n}.c2.{c}.b"], sym_taps(k, 1), co) + cur
                else:
                    cur = conv_cl(lrelu(cur, 0.1), blobs[f"dec.rb.{n}.c.{c}.w"], blobs[f"dec.rb.{n}.c.{c}.b"], sym_taps(k, d), co) + cur
            xs = cur if xs is None else xs + cur
        x = (xs / nk).astype(np.float32)
        ch = co
    xin = lrelu(x, 0.01)
    w = blobs["dec.post_w"]          # [7][C]
    L = xin.shape[0]
    acc = np.zeros((L,), np.float64)
    for k in range(7):
        off = k - 3
        lo, hi = max(0, -off), min(L, L - off)
        acc[lo:hi] += xin[lo + off:hi + off].astype(np.float64) @ w[k].astype(np.float64)
    return np.tanh(acc).astype(np.float32)
