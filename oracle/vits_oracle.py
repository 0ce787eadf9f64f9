"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the phoneme-ids -> audio hot path.

A from-scratch, functional restatement of the arithmetic that the reference freezes into its
ONNX graph, i.e. ``SynthesizerTrn.infer`` (phoonnx_train/vits/models.py:681-722) as wrapped
by ``infer_forward`` (phoonnx_train/export_onnx.py:250-278).  It consumes the *canonical*
weight dict produced by ``phoonnx_b200.weights`` (the very tensors the engine uploads), so
the CUDA path and the oracle see identical inputs.  One utterance at a time: the engine
implements the B=1 semantics that TTSVoice uses (voice.py:350-351).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / --impl
reference leg may import this module.  The product path never does.

Parity pin: ``tests/test_oracle_vs_reference.py`` checks every stage tensor of this file
against the reference's own PyTorch modules (imported from /root/reference in the build
container) and ``tests/golden/*.npz`` holds outputs minted from the reference
(``oracle/make_golden.py``) that are re-checked wherever the reference is absent.

Each function cites the reference lines it follows.  Layout: channel-last ``[T, C]``
numpy arrays at the interface, torch ``[1, C, T]`` inside (CPU fp32, torch ops = MKL/oneDNN,
which is also what makes this a fair multi-threaded CPU baseline).
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F


def _t(W, name):
    return torch.from_numpy(np.ascontiguousarray(W[name], dtype=np.float32))


class VitsOracle:
    def __init__(self, W: Dict[str, np.ndarray], arch):
        self.W = {k: torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)) for k, v in W.items()}
        self.a = arch

    # ------------------------------------------------------------------ small helpers
    def _conv(self, x, name, dilation=1, padding=0, groups=1):
        return F.conv1d(x, self.W[name + ".weight"], self.W.get(name + ".bias"), dilation=dilation,
                        padding=padding, groups=groups)

    def _ln(self, x, name):
        """modules.py:23-26: LayerNorm over channels, eps 1e-5."""
        C = x.shape[1]
        y = F.layer_norm(x.transpose(1, 2), (C,), self.W[name + ".gamma"], self.W[name + ".beta"], 1e-5)
        return y.transpose(1, 2)

    # ------------------------------------------------------------------ text encoder
    def _rel_attention(self, x, li):
        """attentions.py:215-272 restated per SURVEY.md A1: the pad/reshape skew of the
        reference equals a +-window band: scores[i,j] += q_i.E_k[j-i+w], out_i += p_ij E_v[j-i+w]."""
        a = self.a
        pre = f"enc_p.encoder.attn_layers.{li}"
        T = x.shape[2]
        nh, dk, w = a.n_heads, a.k_channels, a.window
        q = self._conv(x, pre + ".conv_q").view(nh, dk, T).transpose(1, 2)  # [h, T, dk]
        k = self._conv(x, pre + ".conv_k").view(nh, dk, T).transpose(1, 2)
        v = self._conv(x, pre + ".conv_v").view(nh, dk, T).transpose(1, 2)
        qs = q / math.sqrt(dk)
        scores = qs @ k.transpose(1, 2)                                     # [h, T, T]
        Ek = self.W[pre + ".emb_rel_k"][0]                                  # [2w+1, dk]
        Ev = self.W[pre + ".emb_rel_v"][0]
        rel = qs @ Ek.T                                                      # [h, T, 2w+1]
        ii = torch.arange(T)
        for r in range(2 * w + 1):
            jj = ii + (r - w)
            ok = (jj >= 0) & (jj < T)
            scores[:, ii[ok], jj[ok]] += rel[:, ii[ok], r]
        p = torch.softmax(scores, dim=-1)
        out = p @ v                                                          # [h, T, dk]
        for r in range(2 * w + 1):
            jj = ii + (r - w)
            ok = (jj >= 0) & (jj < T)
            out[:, ii[ok], :] += p[:, ii[ok], jj[ok]].unsqueeze(-1) * Ev[r]
        out = out.transpose(1, 2).reshape(1, nh * dk, T)
        return self._conv(out, pre + ".conv_o")

    def text_encoder(self, ids: np.ndarray):
        """models.py:198-209 + attentions.py:60-74 (post-LN encoder), attentions.py:386-407 (FFN)."""
        a = self.a
        x = self.W["enc_p.emb.weight"][torch.from_numpy(np.asarray(ids, dtype=np.int64))] * math.sqrt(a.hidden)
        x = x.T.unsqueeze(0)                                                 # [1, H, T]
        kpad_l, kpad_r = (a.enc_kernel - 1) // 2, a.enc_kernel // 2
        for li in range(a.n_layers):
            y = self._rel_attention(x, li)
            x = self._ln(x + y, f"enc_p.encoder.norm_layers_1.{li}")
            h = self._conv(F.pad(x, (kpad_l, kpad_r)), f"enc_p.encoder.ffn_layers.{li}.conv_1")
            h = torch.relu(h)
            h = self._conv(F.pad(h, (kpad_l, kpad_r)), f"enc_p.encoder.ffn_layers.{li}.conv_2")
            x = self._ln(x + h, f"enc_p.encoder.norm_layers_2.{li}")
        stats = self._conv(x, "enc_p.proj")
        m_p, logs_p = stats[:, : a.inter], stats[:, a.inter:]
        return x, m_p, logs_p

    # ------------------------------------------------------------------ duration predictor
    def _dds(self, x, pre):
        """modules.py:117-129 (mask == 1 for a single un-padded utterance)."""
        a = self.a
        for i in range(a.dds_layers):
            d = a.dp_kernel ** i
            pad = (a.dp_kernel * d - d) // 2
            y = self._conv(x, f"{pre}.convs_sep.{i}", dilation=d, padding=pad, groups=x.shape[1])
            y = F.gelu(self._ln(y, f"{pre}.norms_1.{i}"))
            y = self._conv(y, f"{pre}.convs_1x1.{i}")
            y = F.gelu(self._ln(y, f"{pre}.norms_2.{i}"))
            x = x + y
        return x

    @staticmethod
    def spline_inverse(y, h, filter_channels, num_bins=10, tail_bound=5.0):
        """transforms.py:50-191 (inverse, linear tails) restated per element, vectorised
        (SURVEY.md A8).  y: [T], h: [T, 3*num_bins-1]."""
        K = num_bins
        s = 1.0 / math.sqrt(filter_channels)
        eps = 1e-3
        uw, uh, ud = h[:, :K] * s, h[:, K:2 * K] * s, h[:, 2 * K:]
        inside = (y >= -tail_bound) & (y <= tail_bound)
        const = math.log(math.exp(1 - eps) - 1)
        ud = F.pad(ud, (1, 1), value=const)
        widths = eps + (1 - eps * K) * torch.softmax(uw, dim=-1)
        cw = F.pad(torch.cumsum(widths, dim=-1), (1, 0))
        cw = 2 * tail_bound * cw - tail_bound
        cw[:, 0], cw[:, -1] = -tail_bound, tail_bound
        widths = cw[:, 1:] - cw[:, :-1]
        deriv = eps + F.softplus(ud)
        heights = eps + (1 - eps * K) * torch.softmax(uh, dim=-1)
        ch = F.pad(torch.cumsum(heights, dim=-1), (1, 0))
        ch = 2 * tail_bound * ch - tail_bound
        ch[:, 0], ch[:, -1] = -tail_bound, tail_bound
        heights = ch[:, 1:] - ch[:, :-1]
        loc = ch.clone()
        loc[:, -1] += 1e-6                                    # transforms.py:44-47
        yy = torch.where(inside, y, torch.zeros_like(y))
        b_idx = ((yy[:, None] >= loc).sum(-1) - 1).clamp(0, K - 1)[:, None]
        g = lambda t: t.gather(-1, b_idx)[:, 0]               # noqa: E731
        in_cw, in_w, in_ch, in_h = g(cw), g(widths), g(ch), g(heights)
        in_delta = g(heights / widths)
        in_d, in_d1 = g(deriv), g(deriv[:, 1:])
        u = yy - in_ch
        t2 = in_d + in_d1 - 2 * in_delta
        qa = u * t2 + in_h * (in_delta - in_d)
        qb = in_h * in_d - u * t2
        qc = -in_delta * u
        disc = qb.pow(2) - 4 * qa * qc
        root = (2 * qc) / (-qb - torch.sqrt(disc))
        out = root * in_w + in_cw
        return torch.where(inside, out, y)

    def sdp_reverse(self, x, noise_dp: np.ndarray, noise_w: float, g=None):
        """models.py:63-71,108-117: logw = reverse flows([Flip,CF7,Flip,CF5,Flip,CF3,Flip,EA0])."""
        a = self.a
        h = self._conv(x, "dp.pre")
        if g is not None:
            h = h + self._conv(g, "dp.cond")
        h = self._dds(h, "dp.convs")
        cond = self._conv(h, "dp.proj")
        z = torch.from_numpy(np.asarray(noise_dp, dtype=np.float32))[None] * noise_w   # [1,2,T]
        for fi in a.cflows:
            z = torch.flip(z, [1])                                           # modules.py:386
            x0, x1 = z[:, :1], z[:, 1:]
            hh = self._conv(x0, f"dp.flows.{fi}.pre")
            hh = self._dds(hh + cond, f"dp.flows.{fi}.convs")                # modules.py:118-119
            hh = self._conv(hh, f"dp.flows.{fi}.proj")                       # [1, 29, T]
            x1n = self.spline_inverse(x1[0, 0], hh[0].T, a.dp_filter, a.num_bins)
            z = torch.cat([x0, x1n[None, None]], 1)
        z = torch.flip(z, [1])
        m, logs = self.W["dp.flows.0.m"], self.W["dp.flows.0.logs"]
        z = (z - m) * torch.exp(-logs)                                       # modules.py:408
        return z[0, 0]

    def dp_deterministic(self, x, g=None):
        """models.py:151-165 (use_sdp=False)."""
        if g is not None:
            x = x + self._conv(g, "dp.cond")
        pad = self.a.dp_kernel // 2
        h = self._ln(torch.relu(self._conv(x, "dp.conv_1", padding=pad)), "dp.norm_1")
        h = self._ln(torch.relu(self._conv(h, "dp.conv_2", padding=pad)), "dp.norm_2")
        return self._conv(h, "dp.proj")[0, 0]

    # ------------------------------------------------------------------ length regulation
    @staticmethod
    def durations_from_logw(logw: torch.Tensor, length_scale: float):
        """models.py:702-704: w = exp(logw)*mask*length_scale; ceil; y_len = max(sum,1)."""
        w = torch.exp(logw) * 1.0 * length_scale
        w_ceil = torch.ceil(w)
        y_len = int(torch.clamp_min(w_ceil.sum(), 1).long())
        return w_ceil.to(torch.int64), y_len

    @staticmethod
    def frame_index(dur: torch.Tensor, y_len: int):
        """commons.py:116-129 + models.py:711-716 restated: frame j belongs to
        the id whose cumulative duration first exceeds j; frames past sum(dur) (only possible
        when every duration is 0 and y_len was clamped to 1) select nothing (-1)."""
        cum = torch.cumsum(dur, 0)
        j = torch.arange(y_len)
        idx = torch.searchsorted(cum, j, right=True)
        idx[j >= cum[-1]] = -1
        return idx

    # ------------------------------------------------------------------ flow
    def _wn(self, h, pre, g=None):
        """modules.py:184-209."""
        a = self.a
        H = a.hidden
        out = torch.zeros_like(h)
        gc = self._conv(g, pre + ".cond_layer") if g is not None else None
        for i in range(a.wn_layers):
            d = a.wn_dilation_rate ** i
            pad = (a.wn_kernel * d - d) // 2
            xin = self._conv(h, f"{pre}.in_layers.{i}", dilation=d, padding=pad)
            if gc is not None:
                xin = xin + gc[:, i * 2 * H:(i + 1) * 2 * H]
            acts = torch.tanh(xin[:, :H]) * torch.sigmoid(xin[:, H:])       # commons.py:99-106
            rs = self._conv(acts, f"{pre}.res_skip_layers.{i}")
            if i < a.wn_layers - 1:
                h = h + rs[:, :H]
                out = out + rs[:, H:]
            else:
                out = out + rs
        return out

    def flow_reverse(self, z_p, g=None):
        """models.py:247-254 + modules.py:447-466 (mean_only coupling, reverse)."""
        a = self.a
        half = a.inter // 2
        x = z_p
        for fi in a.flow_layers:
            x = torch.flip(x, [1])
            x0, x1 = x[:, :half], x[:, half:]
            h = self._conv(x0, f"flow.flows.{fi}.pre")
            h = self._wn(h, f"flow.flows.{fi}.enc", g)
            m = self._conv(h, f"flow.flows.{fi}.post")
            x = torch.cat([x0, x1 - m], 1)
        return x

    # ------------------------------------------------------------------ decoder
    def decoder(self, z, g=None):
        """models.py:348-368, modules.py:301-314,355-364 (no mask inside dec)."""
        a = self.a
        x = self._conv(z, "dec.conv_pre", padding=3)
        if g is not None:
            x = x + self._conv(g, "dec.cond")
        nk = len(a.rb_kernels)
        for i, (u, k) in enumerate(zip(a.up_rates, a.up_kernels)):
            x = F.leaky_relu(x, 0.1)
            x = F.conv_transpose1d(x, self.W[f"dec.ups.{i}.weight"], self.W[f"dec.ups.{i}.bias"],
                                   stride=u, padding=(k - u) // 2)
            xs = None
            for j in range(nk):
                r = self._resblock(x, i * nk + j, a.rb_kernels[j], a.rb_dilations[j])
                xs = r if xs is None else xs + r
            x = xs / nk
        x = F.leaky_relu(x)                                                  # slope 0.01, models.py:364
        x = F.conv1d(x, self.W["dec.conv_post.weight"], None, padding=3)
        return torch.tanh(x)

    def _resblock(self, x, n, k, dil):
        pre = f"dec.resblocks.{n}"
        if self.a.resblock == "1":
            for c, d in enumerate(dil):
                xt = self._conv(F.leaky_relu(x, 0.1), f"{pre}.convs1.{c}", dilation=d, padding=(k * d - d) // 2)
                xt = self._conv(F.leaky_relu(xt, 0.1), f"{pre}.convs2.{c}", padding=(k - 1) // 2)
                x = xt + x
        else:
            for c, d in enumerate(dil):
                xt = self._conv(F.leaky_relu(x, 0.1), f"{pre}.convs.{c}", dilation=d, padding=(k * d - d) // 2)
                x = xt + x
        return x

    # ------------------------------------------------------------------ whole path
    @torch.no_grad()
    def infer(self, ids: np.ndarray, scales=(0.667, 1.0, 0.8), sid: Optional[int] = None,
              noise_dp: Optional[np.ndarray] = None, noise_z: Optional[np.ndarray] = None,
              logw_override: Optional[np.ndarray] = None, stages: bool = True) -> Dict[str, np.ndarray]:
        """One utterance.  noise_dp: [2, T]; noise_z: [C, >=Ty] (reference layouts,
        models.py:111,718).  Missing noise with a non-zero scale is an error: the oracle is
        deterministic by construction."""
        a = self.a
        noise_scale, length_scale, noise_w = (float(s) for s in scales)
        ids = np.asarray(ids, dtype=np.int64)
        T = ids.shape[0]
        if ids.min() < 0 or ids.max() >= a.n_vocab:
            raise ValueError("phoneme id out of range")
        g = None
        if a.n_speakers > 1:
            if sid is None or not (0 <= int(sid) < a.n_speakers):
                raise ValueError("missing / out-of-range speaker id")
            g = self.W["emb_g.weight"][int(sid)][None, :, None]              # models.py:694
        x, m_p, logs_p = self.text_encoder(ids)
        if logw_override is not None:
            logw = torch.from_numpy(np.asarray(logw_override, dtype=np.float32))
        elif a.use_sdp:
            if noise_dp is None:
                if noise_w != 0.0:
                    raise ValueError("noise_dp required when noise_w != 0")
                noise_dp = np.zeros((2, T), np.float32)
            logw = self.sdp_reverse(x, noise_dp, noise_w, g)
        else:
            logw = self.dp_deterministic(x, g)
        dur, y_len = self.durations_from_logw(logw, length_scale)
        idx = self.frame_index(dur, y_len)
        sel = idx.clamp_min(0)
        valid = (idx >= 0).to(torch.float32)
        m_e = m_p[:, :, sel] * valid                                          # models.py:711-716
        logs_e = logs_p[:, :, sel] * valid
        if noise_z is None:
            if noise_scale != 0.0:
                raise ValueError("noise_z required when noise_scale != 0")
            eps = torch.zeros_like(m_e)
        else:
            eps = torch.from_numpy(np.asarray(noise_z, dtype=np.float32))[None, :, :y_len]
        z_p = m_e + eps * torch.exp(logs_e) * noise_scale                     # models.py:718
        z = self.flow_reverse(z_p, g)
        o = self.decoder(z, g)                                                # models.py:720
        res = {"audio": o[0, 0].numpy(), "durations": dur.numpy().astype(np.int32),
               "y_len": np.int64(y_len)}
        if stages:
            res.update({
                "x": x[0].T.contiguous().numpy(), "m_p": m_p[0].T.contiguous().numpy(),
                "logs_p": logs_p[0].T.contiguous().numpy(), "logw": logw.numpy(),
                "frame_index": idx.numpy().astype(np.int32),
                "z_p": z_p[0].T.contiguous().numpy(), "z": z[0].T.contiguous().numpy(),
            })
        return res


def postprocess_int16(audio: np.ndarray, volume: float = 1.0, normalize: bool = True) -> np.ndarray:
    """voice.py:271-282 + AudioChunk.audio_int16_array voice.py:88-91 (next-row 8f-1)."""
    audio = np.asarray(audio, dtype=np.float32)
    if normalize:
        max_val = np.max(np.abs(audio)) if audio.size else np.float32(0)
        if max_val < 1e-8:
            audio = np.zeros_like(audio)
        else:
            audio = audio / max_val
    if volume != 1.0:
        audio = audio * np.float32(volume)
    audio = np.clip(audio, -1.0, 1.0).astype(np.float32)
    return np.clip(audio * np.float32(32767.0), -32767.0, 32767.0).astype(np.int16)
