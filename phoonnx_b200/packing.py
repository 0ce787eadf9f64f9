"""Canonical tensors -> kernel-layout blobs (the "weight loader -> device buffers" half).

Input: the canonical dict of ``phoonnx_b200.weights`` (state_dict names of
phoonnx_train/vits/models.py, weight-norm folded).  Output: ``{blob name: ndarray}`` in
the layouts the CUDA kernels index directly, plus scalar options:

  conv   "<n>.w"   fp32 [taps][C_in][N4]     W[tap][ci][n] = weight[n, ci, tap]      (N4 = N rounded up to 4)
         "<n>.b"   fp32 [N4]
         "<n>.wtc" bf16 [taps][C_in/8][N16][8] (tcgen05 K-major no-swizzle chunks)   (N16 = N rounded up to 16)
  Folded at pack time (SURVEY.md A10, 8a rows 9,14,19,22):
    * Flip (modules.py:384-391) of the coupling flow: layers executed at even positions read
      the upper half of the physical tensor with ``pre`` input channels reversed and write the
      lower half with ``post`` output channels reversed;
    * WN in_layer output channels interleaved (tanh_c, sigmoid_c) so the gate
      (commons.py:99-106) is a per-thread epilogue;
    * ConvTranspose1d (k = 2*stride, pad = stride/2) as two 2-tap polyphase groups whose
      output row-block is the contiguous [stride][C_out] span of the channel-last result;
    * speaker conditioning 1x1 convs of ``g = emb_g[sid]`` (models.py:694, modules.py:188-195)
      pre-computed per speaker into bias tables.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

from .weights import VitsArch


def _rup(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """Round-to-nearest-even fp32 -> bf16 (returned as uint16 bit patterns)."""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    rounded = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return rounded.astype(np.uint16)


def bf16_round(x: np.ndarray) -> np.ndarray:
    return (f32_to_bf16_bits(x).astype(np.uint32) << 16).view(np.float32)


SPLIT3_MAX_CIN = 96    # widest K slice whose hi+lo activation planes, double-buffered, fit one CTA's shared memory next to a weight ring.
                       # (r01 A/B: 48-wide slices deepen the ring from 2-3 to 5 stages but double the activation-tile hand-offs, each a
                       # ~5k-cycle load -> convert round trip: text stage 74.3 -> 78.3 ms.  The text GEMMs are bound by shared-memory
                       # capacity: A hi/lo planes + weight ring + epilogue staging)


def split3_slice(cin: int) -> int:
    """Input channels per K slice of a bf16x3 convolution: the largest divisor of ``cin`` that is a multiple
    of 16 and <= SPLIT3_MAX_CIN (0: shape unsupported).  Must match ``mkconv`` in csrc/engine.cu."""
    for s in range(min(cin, SPLIT3_MAX_CIN), 15, -1):
        if cin % s == 0 and s % 16 == 0:
            return s
    return 0


def tc_layout(w_tcn16: np.ndarray) -> np.ndarray:
    """[tap][ci][n16] fp32 -> tcgen05 K-major no-swizzle chunks [tap][ci/8][n16][8] as bf16 bits."""
    taps, cin, n16 = w_tcn16.shape
    wt = w_tcn16.reshape(taps, cin // 8, 8, n16).transpose(0, 1, 3, 2)
    return f32_to_bf16_bits(np.ascontiguousarray(wt))


def pack_conv(out: Dict[str, np.ndarray], name: str, w_tcn: np.ndarray, bias, tc: bool = False,
              tc3: bool = False) -> None:
    """w_tcn: [taps, C_in, N] (already in kernel tap order).

    tc3: additionally emit the fp32-faithful bf16x3 operands "<n>.wtc3.<j>" (one per K slice j): per tap the
    input channels are [wh | wl | wh] with wh = bf16(w), wl = bf16(w - wh); the kernel multiplies them with
    [xh | xh | xl] (csrc/conv_tc.cuh, ConvArgs::split3)."""
    taps, cin, n = w_tcn.shape
    n4, n16 = _rup(n, 4), _rup(n, 16)
    w = np.zeros((taps, cin, n4), np.float32)
    w[:, :, :n] = w_tcn
    out[name + ".w"] = w
    if bias is not None:
        b = np.zeros((n4,), np.float32)
        b[:n] = bias
        out[name + ".b"] = b
    if tc and cin % 16 == 0 and n % 16 == 0:
        wt = np.zeros((taps, cin, n16), np.float32)
        wt[:, :, :n] = w_tcn
        out[name + ".wtc"] = tc_layout(wt)
    if tc3 and cin % 16 == 0 and n % 16 == 0 and split3_slice(cin):
        sl = split3_slice(cin)
        hi = bf16_round(np.asarray(w_tcn, np.float32))
        lo = bf16_round(np.asarray(w_tcn, np.float32) - hi)
        for j in range(cin // sl):
            seg = slice(j * sl, (j + 1) * sl)
            out[f"{name}.wtc3.{j}"] = tc_layout(np.ascontiguousarray(
                np.concatenate([hi[:, seg], lo[:, seg], hi[:, seg]], axis=1)))


def _conv_std(W, name) -> Tuple[np.ndarray, np.ndarray]:
    w = W[name + ".weight"]                       # [N, C_in, k]
    return np.ascontiguousarray(w.transpose(2, 1, 0)), W.get(name + ".bias")


def pack_model(W: Dict[str, np.ndarray], a: VitsArch, tc: bool = True):
    """Returns (blobs, options)."""
    o: Dict[str, np.ndarray] = {}
    opts: Dict[str, float] = {}
    H, C = a.hidden, a.inter
    o["enc.emb"] = np.ascontiguousarray(W["enc_p.emb.weight"], np.float32)
    for i in range(a.n_layers):
        p = f"enc_p.encoder.attn_layers.{i}"
        wq, bq = _conv_std(W, p + ".conv_q")
        wk, bk = _conv_std(W, p + ".conv_k")
        wv, bv = _conv_std(W, p + ".conv_v")
        pack_conv(o, f"enc.{i}.qkv", np.concatenate([wq, wk, wv], axis=2), np.concatenate([bq, bk, bv]), tc3=tc)
        o[f"enc.{i}.rel_k"] = np.ascontiguousarray(W[p + ".emb_rel_k"][0], np.float32)
        o[f"enc.{i}.rel_v"] = np.ascontiguousarray(W[p + ".emb_rel_v"][0], np.float32)
        pack_conv(o, f"enc.{i}.o", *_conv_std(W, p + ".conv_o"), tc3=tc)
        pack_conv(o, f"enc.{i}.ffn1", *_conv_std(W, f"enc_p.encoder.ffn_layers.{i}.conv_1"), tc3=tc)
        pack_conv(o, f"enc.{i}.ffn2", *_conv_std(W, f"enc_p.encoder.ffn_layers.{i}.conv_2"), tc3=tc)
        for j in (1, 2):
            o[f"enc.{i}.ln{j}.g"] = W[f"enc_p.encoder.norm_layers_{j}.{i}.gamma"]
            o[f"enc.{i}.ln{j}.b"] = W[f"enc_p.encoder.norm_layers_{j}.{i}.beta"]
    pack_conv(o, "enc.proj", *_conv_std(W, "enc_p.proj"), tc3=tc)

    emb_g = W.get("emb_g.weight") if a.n_speakers > 1 else None

    def cond_table(name):     # 1x1 conv of g = emb_g[s]: [n_spk, N]
        w = W[name + ".weight"][:, :, 0].astype(np.float64)     # [N, gin]
        return (emb_g.astype(np.float64) @ w.T + W[name + ".bias"].astype(np.float64)).astype(np.float32)

    def pack_dds(dst, src):
        for i in range(a.dds_layers):
            o[f"{dst}.{i}.dw_w"] = np.ascontiguousarray(W[f"{src}.convs_sep.{i}.weight"][:, 0, :].T, np.float32)  # [k][C]
            o[f"{dst}.{i}.dw_b"] = W[f"{src}.convs_sep.{i}.bias"]
            pack_conv(o, f"{dst}.{i}.pw", *_conv_std(W, f"{src}.convs_1x1.{i}"), tc3=tc)
            o[f"{dst}.{i}.ln1.g"] = W[f"{src}.norms_1.{i}.gamma"]
            o[f"{dst}.{i}.ln1.b"] = W[f"{src}.norms_1.{i}.beta"]
            o[f"{dst}.{i}.ln2.g"] = W[f"{src}.norms_2.{i}.gamma"]
            o[f"{dst}.{i}.ln2.b"] = W[f"{src}.norms_2.{i}.beta"]

    if a.use_sdp:
        pack_conv(o, "dp.pre", *_conv_std(W, "dp.pre"), tc3=tc)
        pack_conv(o, "dp.proj", *_conv_std(W, "dp.proj"), tc3=tc)
        pack_dds("dp.convs", "dp.convs")
        for fi in a.cflows:
            o[f"dp.flows.{fi}.pre_w"] = np.ascontiguousarray(W[f"dp.flows.{fi}.pre.weight"][:, 0, 0], np.float32)
            o[f"dp.flows.{fi}.pre_b"] = W[f"dp.flows.{fi}.pre.bias"]
            pack_dds(f"dp.flows.{fi}.convs", f"dp.flows.{fi}.convs")
            # 3 * num_bins - 1 = 29 spline parameters: padded with zero output channels to a multiple of 16 so that the projection
            # runs on the tensor cores like the rest of the text side (bf16x3); k_spline_inverse never reads the padding
            pw, pb = _conv_std(W, f"dp.flows.{fi}.proj")
            n16 = _rup(pw.shape[2], 16)
            pwp = np.zeros(pw.shape[:2] + (n16,), np.float32); pwp[:, :, :pw.shape[2]] = pw
            pbp = np.zeros((n16,), np.float32); pbp[:pw.shape[2]] = pb
            pack_conv(o, f"dp.flows.{fi}.proj", pwp, pbp, tc3=tc)
        # the final Flip puts logical channel 0 on logw; ElementwiseAffine indexes logical channels (modules.py:408)
        opts["dp.ea_m"] = float(W["dp.flows.0.m"][0, 0])
        opts["dp.ea_logs"] = float(W["dp.flows.0.logs"][0, 0])
    else:
        pack_conv(o, "dp.conv_1", *_conv_std(W, "dp.conv_1"), tc3=tc)
        pack_conv(o, "dp.conv_2", *_conv_std(W, "dp.conv_2"), tc3=tc)
        pack_conv(o, "dp.proj", *_conv_std(W, "dp.proj"))
        for j in (1, 2):
            o[f"dp.norm_{j}.g"] = W[f"dp.norm_{j}.gamma"]
            o[f"dp.norm_{j}.b"] = W[f"dp.norm_{j}.beta"]
    if emb_g is not None:
        o["dp.cond_tab"] = cond_table("dp.cond")
        o["dec.cond_tab"] = cond_table("dec.cond")

    half = C // 2
    inter = np.empty((2 * H,), np.int64)           # interleave (tanh_c, sigmoid_c)
    inter[0::2] = np.arange(H)
    inter[1::2] = np.arange(H) + H
    for s, fi in enumerate(a.flow_layers):
        p = f"flow.flows.{fi}"
        flipped = (s % 2 == 0)
        w, b = _conv_std(W, p + ".pre")             # [1, half, H]
        if flipped:
            w = w[:, ::-1, :]
        pack_conv(o, f"flow.{s}.pre", np.ascontiguousarray(w), b, tc=tc)
        for i in range(a.wn_layers):
            w, b = _conv_std(W, f"{p}.enc.in_layers.{i}")      # [k, H, 2H]
            pack_conv(o, f"flow.{s}.in.{i}", np.ascontiguousarray(w[:, :, inter]), b[inter], tc=tc)
            w, b = _conv_std(W, f"{p}.enc.res_skip_layers.{i}")
            pack_conv(o, f"flow.{s}.rs.{i}", w, b, tc=tc)
        if emb_g is not None:
            tab = cond_table(p + ".enc.cond_layer")              # [n_spk, 2H * layers]
            for i in range(a.wn_layers):
                o[f"flow.{s}.cond_tab.{i}"] = np.ascontiguousarray(tab[:, i * 2 * H:(i + 1) * 2 * H][:, inter])
        w, b = _conv_std(W, p + ".post")            # [1, H, half]
        if flipped:
            w, b = w[:, :, ::-1], b[::-1]
        pack_conv(o, f"flow.{s}.post", np.ascontiguousarray(w), np.ascontiguousarray(b), tc=tc)
        if tc:
            # tensor-core form of the WN tail (modules.py:196-209 + 447-466): the skip path is linear, so
            #   m = post(sum_l skip_l(acts_l)) = sum_l (W_skip_l @ W_post) acts_l + (sum_l b_skip_l) @ W_post + b_post
            # is ONE GEMM over the concatenated gate outputs of all layers -- no skip accumulator in HBM, no separate
            # post conv; the res_skip convs keep their residual half only
            wp = np.asarray(w[0], np.float64)                                   # [H, half]
            bsum = np.zeros((H,), np.float64)
            for i in range(a.wn_layers):
                wr, br = _conv_std(W, f"{p}.enc.res_skip_layers.{i}")          # [1, H, 2H] (last layer: [1, H, H])
                if i < a.wn_layers - 1:
                    pack_conv(o, f"flow.{s}.rsr.{i}", np.ascontiguousarray(wr[:, :, :H]), np.ascontiguousarray(br[:H]), tc=True)
                    wsk, bsk = wr[0][:, H:], br[H:]
                else:
                    wsk, bsk = wr[0], br
                pack_conv(o, f"flow.{s}.mskip.{i}", (np.asarray(wsk, np.float64) @ wp).astype(np.float32)[None],
                          None, tc=True)
                bsum += np.asarray(bsk, np.float64)
            o[f"flow.{s}.mskip.b"] = (bsum @ wp + np.asarray(b, np.float64)).astype(np.float32)

    pack_conv(o, "dec.pre", *_conv_std(W, "dec.conv_pre"), tc=tc)
    nk = len(a.rb_kernels)
    for i, (u, k) in enumerate(zip(a.up_rates, a.up_kernels)):
        w = W[f"dec.ups.{i}.weight"]                # [C_in, C_out, k] (ConvTranspose layout)
        b = W[f"dec.ups.{i}.bias"]
        cin, cout, _ = w.shape
        pad = (k - u) // 2
        assert k == 2 * u and u % 2 == 0 and pad == u // 2
        hu = u // 2
        wa = np.zeros((2, cin, hu * cout), np.float32)   # taps (-1, 0), phases [0, u/2)
        wb = np.zeros((2, cin, hu * cout), np.float32)   # taps (0, +1), phases [u/2, u)
        for p_ in range(hu):
            r = p_ + pad                                   # < u
            wa[0, :, p_ * cout:(p_ + 1) * cout] = w[:, :, r + u]   # t_in = q - 1
            wa[1, :, p_ * cout:(p_ + 1) * cout] = w[:, :, r]       # t_in = q
            r2 = (p_ + hu) + pad                           # >= u
            wb[0, :, p_ * cout:(p_ + 1) * cout] = w[:, :, r2]      # t_in = q
            wb[1, :, p_ * cout:(p_ + 1) * cout] = w[:, :, r2 - u]  # t_in = q + 1
        pack_conv(o, f"dec.ups.{i}.A", wa, np.tile(b, hu), tc=tc)
        pack_conv(o, f"dec.ups.{i}.B", wb, np.tile(b, hu), tc=tc)
        for j in range(nk):
            n = i * nk + j
            for c in range(len(a.rb_dilations[j])):
                if a.resblock == "1":
                    pack_conv(o, f"dec.rb.{n}.c1.{c}", *_conv_std(W, f"dec.resblocks.{n}.convs1.{c}"), tc=tc)
                    pack_conv(o, f"dec.rb.{n}.c2.{c}", *_conv_std(W, f"dec.resblocks.{n}.convs2.{c}"), tc=tc)
                else:
                    pack_conv(o, f"dec.rb.{n}.c.{c}", *_conv_std(W, f"dec.resblocks.{n}.convs.{c}"), tc=tc)
    o["dec.post_w"] = np.ascontiguousarray(W["dec.conv_post.weight"][0].T, np.float32)     # [7][C]
    return {k: np.ascontiguousarray(v) for k, v in o.items()}, opts
