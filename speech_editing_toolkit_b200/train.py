"""Training-mode DiffNet behind the C ABI (BASELINE configs[4]).

`DiffNetB200.forward` under autograd routes here: the forward and the activation-gradient chain of the backward run natively
(`fse_train_forward` / `fse_train_backward`, csrc/denoiser_train.cuh: every data-path GEMM is a conv-GEMM launch of the library, tcgen05
in the tensor-core modes), and so do the weight gradients — GEMMs `dW = dY^T A` with K = all frames over tensors the native code leaves
in its workspace, taken by `fse_wgrad` (csrc/wgrad.cu: tcgen05 with MN-major operands, no transposed copies; `FSE_TRAIN_WGRAD=lib` and
the fp32 checking mode use torch.mm = cuBLASLt instead, the cross-check).  The bias sums and the timestep-MLP path (a [B, 256] problem)
stay torch reductions / a torch graph; the losses are native (`fse_mel_loss_*`); optimizer and the NCCL gradient all-reduce are torch
plumbing (`train_step`, `BucketedAllReduce`).

Reference semantics: modules/speech_editing/spec_denoiser/diffnet.py:60-132 under torch.autograd; the call site is
spec_denoiser.py:168-176 (`x_0_pred = self.denoise_fn(x_t, t, cond) * nonpadding`).
"""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import MODES, check
from .engine import _Workspace, _need_cuda, _ptr, _stream

N_BUFS = 18


class DiffNetTrainer:
    """ctypes owner of an `fse_trainer` handle."""

    def __init__(self, n_mels=80, hidden=192, channels=256, layers=20, dilation_cycle_length=1, mode="tc_bf16"):
        self.cfg = _lib.DenoiserConfig(n_mels, hidden, channels, layers, dilation_cycle_length, MODES[mode])
        self.mode = mode
        self.es = 2 if mode in ("tc_bf16", "simt_bf16") else 4
        self.op_dtype = torch.bfloat16 if self.es == 2 else torch.float32
        self._h = C.c_void_p()
        check(_lib.lib().fse_train_create(C.byref(self.cfg), C.byref(self._h)))
        self._ws = _Workspace()
        self._keep = None
        self._shape = None

    def __del__(self):
        try:
            if getattr(self, "_h", None) and self._h.value:
                _lib.lib().fse_train_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def load(self, named: Dict[str, torch.Tensor]):
        """`named`: the reference's `denoise_fn.*` state_dict keys -> fp32 CUDA tensors (the live parameters); call once per step."""
        keep = []
        arr = (_lib.Tensor * len(named))()
        for i, (k, v) in enumerate(named.items()):
            _need_cuda(v)
            t = v.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            name = k.encode()
            keep.append((name, t))
            arr[i].name = name
            arr[i].data = C.cast(C.c_void_p(t.data_ptr()), C.POINTER(C.c_float))
            arr[i].numel = t.numel()
        check(_lib.lib().fse_train_load_weights_device(self._h, arr, len(named), _stream()))
        self._keep = keep            # the library reads the bias tensors again in forward / backward

    def _workspace(self, B, T, device):
        nbytes = _lib.lib().fse_train_workspace_bytes(self._h, B, T)
        return self._ws.get(nbytes, device)

    def forward(self, x_t: torch.Tensor, cond_bth: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
        _need_cuda(x_t, cond_bth, d)
        B, M, T = x_t.shape
        x_t, cond_bth, d = x_t.contiguous().float(), cond_bth.contiguous().float(), d.contiguous().float()
        assert cond_bth.shape == (B, T, self.cfg.hidden) and d.shape == (self.cfg.layers, B, self.cfg.channels)
        x0 = torch.empty_like(x_t)
        ws, nbytes = self._workspace(B, T, x_t.device)
        check(_lib.lib().fse_train_forward(self._h, _ptr(x_t), _ptr(cond_bth), _ptr(d), _ptr(x0), B, T, ws, nbytes, _stream()))
        self._shape = (B, T, ws, nbytes, cond_bth)
        return x0

    def backward(self, dx0: torch.Tensor) -> torch.Tensor:
        B, T, ws, nbytes, cond = self._shape
        dx0 = dx0.contiguous().float()
        dcond = torch.empty(B, T, self.cfg.hidden, dtype=torch.float32, device=dx0.device)
        check(_lib.lib().fse_train_backward(self._h, _ptr(dx0), _ptr(dcond), B, T, ws, nbytes, _stream()))
        return dcond

    def views(self) -> Dict[str, torch.Tensor]:
        """Typed views of the workspace buffers the weight gradients are made of (valid until the next forward)."""
        B, T, ws, nbytes, cond = self._shape
        off = (C.c_int64 * N_BUFS)()
        check(_lib.lib().fse_train_layout(self._h, B, T, off, N_BUFS))
        buf = self._ws.buf
        base = ws.value - buf.data_ptr()
        Cc, H, M, L, N = self.cfg.channels, self.cfg.hidden, self.cfg.n_mels, self.cfg.layers, B * T

        def v(i, shape, dtype):
            n = math.prod(shape) * (2 if dtype == torch.bfloat16 else 4)
            return buf[base + off[i]: base + off[i] + n].view(dtype).view(*shape)
        op = self.op_dtype
        return {"x_rows": v(0, (N, M), op), "h0": v(1, (N, Cc), op), "hin": v(5, (L, B, T, Cc), op), "u": v(8, (L, N, Cc), op),
                "s": v(9, (N, Cc), op), "r": v(10, (N, Cc), op), "cond": (v(11, (B, T, H), op) if self.es == 2 else cond),
                "dx_rows": v(12, (N, M), op), "dz": v(13, (N, Cc), op), "dS": v(14, (N, Cc), op), "dh0": v(15, (N, Cc), torch.float32),
                "dres": v(16, (L, N, Cc), op), "dy": v(17, (B, T, L * 2 * Cc), op)}


class WeightGradGemm:
    """`fse_wgrad` (csrc/wgrad.cu): Out[m, n, j] = sum_{b,t} P[b, t, m] Q[b, t + offs[j], n] on tcgen05 with MN-major operands, straight
    from the [B, T, channels] buffers of the native backward (strided views included).  One workspace per device, shared by the calls of
    a backward pass (its 4 KB head of arrival counters is zero between calls)."""
    _ws: Dict[int, torch.Tensor] = {}

    def __init__(self, mode: str):
        self.mode = MODES[mode]
        self.es = 2 if mode == "tc_bf16" else 4

    def _workspace(self, nbytes: int, device) -> torch.Tensor:
        key = device.index if device.index is not None else torch.cuda.current_device()
        ws = WeightGradGemm._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = torch.zeros(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
            WeightGradGemm._ws[key] = ws
        return ws

    def __call__(self, P: torch.Tensor, Q: torch.Tensor, out: torch.Tensor, offs=(0,)):
        """P [B, T, M] and Q [B, T, N]: views with unit channel stride and a common (B, T) grid; out fp32 [M, N] or [M, N, taps] (any strides)."""
        return self.group([(P, Q, out, offs)])[0]

    def group(self, problems):
        """Up to 4 GEMMs (P, Q, out, offs) over the same (B, T) grid in one launch (fse_wgrad_group); returns the outputs."""
        B, T = problems[0][0].shape[:2]
        arr = (_lib.WgradProblem * len(problems))()
        keep = []
        for i, (P, Q, out, offs) in enumerate(problems):
            M, N, taps = P.shape[2], Q.shape[2], len(offs)
            assert tuple(P.shape[:2]) == (B, T) and tuple(Q.shape[:2]) == (B, T) and P.stride(2) == 1 and Q.stride(2) == 1
            assert B == 1 or (P.stride(0) == T * P.stride(1) and Q.stride(0) == T * Q.stride(1)), "utterances must follow each other at the row pitch"
            assert out.dtype == torch.float32 and tuple(out.shape[:2]) == (M, N) and (out.dim() == 2 or out.shape[2] == taps)
            o = (C.c_int32 * taps)(*[int(v) for v in offs])
            keep.append(o)
            a = arr[i]
            a.P, a.ldp, a.Q, a.ldq, a.M, a.N = P.data_ptr(), P.stride(1), Q.data_ptr(), Q.stride(1), M, N
            a.offs, a.ntaps, a.out = C.cast(o, C.POINTER(C.c_int32)), taps, out.data_ptr()
            a.ld_m, a.ld_n, a.ld_j = out.stride(0), out.stride(1), out.stride(2) if out.dim() == 3 else 0
        lib = _lib.lib()
        nbytes = lib.fse_wgrad_group_workspace_bytes(self.mode, arr, len(problems), B, T)
        ws = self._workspace(nbytes, problems[0][0].device)
        check(lib.fse_wgrad_group(self.mode, arr, len(problems), B, T, _ptr(ws), ws.numel(), _stream()))
        return [pr[2] for pr in problems]


def param_names(layers: int) -> List[str]:
    """Parameters whose gradients `DiffNetFunction.backward` returns, in this order (the timestep path — mlp.*, diffusion_projection.* —
    enters through `d` and stays a torch graph)."""
    names = ["input_projection.weight", "input_projection.bias"]
    for l in range(layers):
        p = f"residual_layers.{l}."
        names += [p + "dilated_conv.weight", p + "dilated_conv.bias", p + "conditioner_projection.weight", p + "conditioner_projection.bias",
                  p + "output_projection.weight", p + "output_projection.bias"]
    return names + ["skip_projection.weight", "skip_projection.bias", "output_projection.weight", "output_projection.bias"]


class DiffNetFunction(torch.autograd.Function):
    """x0 = DiffNet(x_t, t, cond) with d = diffusion_projection_l(temb) given; gradients for cond, d and the parameters."""

    @staticmethod
    def forward(ctx, trainer: DiffNetTrainer, dil: List[int], hook, x_t, cond_bth, d, *params):
        L = trainer.cfg.layers
        trainer.load(dict(zip(param_names(L), params)))
        x0 = trainer.forward(x_t, cond_bth, d)
        ctx.trainer, ctx.dil, ctx.hook = trainer, dil, hook
        ctx.save_for_backward(*[p for n, p in zip(param_names(L), params) if n.endswith("dilated_conv.weight")])
        return x0

    @staticmethod
    def backward(ctx, dx0):
        tr: DiffNetTrainer = ctx.trainer
        L, Cc = tr.cfg.layers, tr.cfg.channels
        dcond = tr.backward(dx0)
        v = tr.views()
        B, T = v["cond"].shape[:2]
        N = B * T
        f32 = torch.float32
        grads: Dict[str, torch.Tensor] = {}
        hook = ctx.hook                                        # called with (name -> grad) groups as they become ready (all-reduce overlap)

        # Weight gradients: native tcgen05 GEMMs over MN-major operands in the tensor-core modes (fse_wgrad); FSE_TRAIN_WGRAD=lib (and
        # the fp32 checking mode) take them with library GEMMs instead — the cross-check of the native kernel.
        native = tr.mode in ("tc_bf16", "tc_tf32") and dx0.is_cuda and os.environ.get("FSE_TRAIN_WGRAD", "native") != "lib"
        wg = WeightGradGemm(tr.mode) if native else None

        def mm(a, b):                                          # a^T b over all rows: operand-dtype GEMM, fp32 accumulator written out as fp32
            if native and a.dtype == tr.op_dtype and b.dtype == tr.op_dtype:
                out = torch.empty(a.shape[1], b.shape[1], dtype=f32, device=a.device)
                return wg(a.unsqueeze(0), b.unsqueeze(0), out)
            return torch.mm(a.t(), b) if a.dtype == f32 else torch.mm(a.t(), b, out_dtype=f32)
        dx_rows, r, dz, s, dS = v["dx_rows"], v["r"], v["dz"], v["s"], v["dS"]
        g = {"output_projection.weight": mm(dx_rows, r)[:, :, None], "output_projection.bias": dx0.sum((0, 2)),
             "skip_projection.weight": mm(dz, s)[:, :, None], "skip_projection.bias": dz.sum(0, dtype=f32)}
        grads.update(g)
        if hook:
            hook(g)
        # every column sum of the backward in a handful of reductions over the whole [.., L * 2C] buffers (fp32 accumulation of the
        # operand-dtype values, no per-layer cast copies): per-item sums of dy, its first / last `dil` frames, and the sums of dres / dS
        dy_all = v["dy"]                                                                 # [B, T, L * 2C]
        dysum_all = dy_all.sum(1, dtype=f32).view(B, L, 2 * Cc)                          # [B, L, 2C]
        edge = {}
        for dil in sorted(set(ctx.dil)):
            edge[dil] = (dy_all[:, :dil].sum(1, dtype=f32).view(B, L, 2 * Cc), dy_all[:, T - dil:].sum(1, dtype=f32).view(B, L, 2 * Cc))
        dres_sum = v["dres"].sum(1, dtype=f32)                                           # [L, C]
        dS_sum = dS.sum(0, dtype=f32)
        dy2d = dy_all.view(N, L * 2 * Cc)
        cond2d = v["cond"].reshape(N, -1)
        gcond_all = None if native else mm(dy2d, cond2d)                                 # library path: conditioner_projection of every layer as ONE GEMM [L * 2C, H]
        # d_l enters as hin = h + d_l inside the zero padding: sum_t of the conv's input gradient, tap by tap (three batched matmuls)
        w_all = torch.stack([w.to(f32) for w in ctx.saved_tensors])                      # [L, 2C, C, 3]
        head_all = torch.stack([edge[ctx.dil[l]][0][:, l] for l in range(L)])            # [L, B, 2C]: frames the tap with offset -dil never reads
        tail_all = torch.stack([edge[ctx.dil[l]][1][:, l] for l in range(L)])
        dys = dysum_all.transpose(0, 1)                                                  # [L, B, 2C]
        dd = torch.bmm(dys - head_all, w_all[..., 0]) + torch.bmm(dys, w_all[..., 1]) + torch.bmm(dys - tail_all, w_all[..., 2])
        for l in range(L - 1, -1, -1):
            p = f"residual_layers.{l}."
            dil = ctx.dil[l]
            dy = dy2d[:, l * 2 * Cc:(l + 1) * 2 * Cc]                                    # [N, 2C], row stride L * 2C: a GEMM operand as it lies
            dy3 = dy_all[:, :, l * 2 * Cc:(l + 1) * 2 * Cc]
            hin = v["hin"][l]                                                            # [B, T, C]
            hin2d = hin.reshape(N, Cc)
            gw = torch.empty(2 * Cc, Cc, 3, dtype=f32, device=dx0.device)                # y[t] += W_j hin[t + off], off = (j - 1) dil
            gop = torch.empty(2 * Cc, Cc, dtype=f32, device=dx0.device)                  # gradient of o = [res | skip] against u_l
            if native:
                # the four weight gradients of the layer as ONE launch: the three conv taps are Q rows shifted by TMA (zero fill per
                # utterance = the conv's padding), the strided dy view is a tensor map with a wider pitch
                gcond = torch.empty(2 * Cc, cond2d.shape[1], dtype=f32, device=dx0.device)
                u3 = v["u"][l].view(B, T, Cc)
                wg.group([(dy3, hin, gw, (-dil, 0, dil)), (dy3, v["cond"].view(B, T, -1), gcond, (0,)),
                          (v["dres"][l].view(B, T, Cc), u3, gop[:Cc], (0,)), (dS.view(B, T, Cc), u3, gop[Cc:], (0,))])
            else:
                # library GEMMs: the shifted taps over the flat row sequence shifted by `dil` rows (no copies), minus the (B - 1) dil row
                # pairs that straddle two utterances
                gw[:, :, 1] = mm(dy, hin2d)
                gw[:, :, 0] = mm(dy[dil:], hin2d[:N - dil])
                gw[:, :, 2] = mm(dy[:N - dil], hin2d[dil:])
                if B > 1:
                    gw[:, :, 0] -= mm(dy3[1:, :dil].reshape(-1, 2 * Cc), hin[:-1, T - dil:].reshape(-1, Cc))
                    gw[:, :, 2] -= mm(dy3[:-1, T - dil:].reshape(-1, 2 * Cc), hin[1:, :dil].reshape(-1, Cc))
                gcond = gcond_all[l * 2 * Cc:(l + 1) * 2 * Cc]
                gop[:Cc] = mm(v["dres"][l], v["u"][l])
                gop[Cc:] = mm(dS, v["u"][l])
            gb = dysum_all[:, l].sum(0)
            g = {p + "dilated_conv.weight": gw, p + "dilated_conv.bias": gb,
                 p + "conditioner_projection.weight": gcond[:, :, None],
                 p + "conditioner_projection.bias": gb,
                 p + "output_projection.weight": gop[:, :, None], p + "output_projection.bias": torch.cat([dres_sum[l], dS_sum])}
            grads.update(g)
            if hook:
                hook(g)
        dpre = v["dh0"] * (v["h0"] > 0)
        g = {"input_projection.weight": mm(dpre.to(tr.op_dtype), v["x_rows"])[:, :, None], "input_projection.bias": dpre.sum(0)}
        grads.update(g)
        if hook:
            hook(g)
        return (None, None, None, None, dcond, dd) + tuple(grads[n] for n in param_names(L))


def sinusoidal_pos_emb(t: torch.Tensor, dim: int) -> torch.Tensor:
    """diffnet.py:34-46"""
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half, device=t.device) * -emb)
    emb = t[:, None] * emb[None, :]
    return torch.cat((emb.sin(), emb.cos()), dim=-1)


def diffnet_train_forward(module, spec: torch.Tensor, diffusion_step: torch.Tensor, cond: torch.Tensor, grad_hook=None) -> torch.Tensor:
    """DiffNet.forward (diffnet.py:110-132) for a `DiffNetB200` module under autograd: spec [B,1,M,T], diffusion_step [B], cond [B,H,T]."""
    tr = getattr(module, "_trainer", None)
    if tr is None or tr.mode != module.mode:
        tr = DiffNetTrainer(n_mels=module.in_dims, hidden=module.encoder_hidden, channels=module.channels, layers=module.n_layers,
                            dilation_cycle_length=module.dilation_cycle_length, mode=module.mode)
        module._trainer = tr
    e = sinusoidal_pos_emb(diffusion_step.float(), module.channels)                      # the timestep path: a torch graph
    e = module.mlp[0](e)
    e = e * torch.tanh(torch.nn.functional.softplus(e))                                  # Mish (diffnet.py:14-16)
    temb = module.mlp[2](e)
    d = torch.stack([layer.diffusion_projection(temb) for layer in module.residual_layers])   # [L, B, C]
    sd = dict(module.named_parameters())
    params = [sd[n] for n in param_names(module.n_layers)]
    dil = [2 ** (i % module.dilation_cycle_length) for i in range(module.n_layers)]
    x0 = DiffNetFunction.apply(tr, dil, grad_hook, spec[:, 0], cond.transpose(1, 2), d, *params)
    return x0[:, None]


class BucketedAllReduce:
    """Gradient all-reduce overlapped with the rest of the backward: every group of gradients the backward hands over (one per
    residual layer, top layer first) is flattened and all-reduced asynchronously on NCCL while the following layers' weight-gradient
    GEMMs run; `finish()` (after loss.backward()) waits, writes the averages into the parameters' .grad and reduces whatever did not
    come through the hook (the timestep path).  This is what DistributedDataParallel's buckets do for the reference
    (utils/commons/trainer.py:475-479)."""

    def __init__(self, named_params: Dict[str, torch.nn.Parameter], group=None):
        import torch.distributed as dist
        self.dist, self.group, self.pending, self.named, self.seen = dist, group, [], dict(named_params), set()

    def active(self) -> bool:
        return self.dist.is_available() and self.dist.is_initialized() and self.dist.get_world_size(self.group) > 1

    def __call__(self, named_grads: Dict[str, torch.Tensor]):
        if not self.active():
            return
        names = list(named_grads)
        flat = torch.cat([named_grads[n].reshape(-1) for n in names])
        work = self.dist.all_reduce(flat, group=self.group, async_op=True)
        self.pending.append((work, flat, names))
        self.seen.update(names)

    def finish(self):
        if not self.active():
            return
        world = self.dist.get_world_size(self.group)
        for work, flat, names in self.pending:
            work.wait()
            o = 0
            for n in names:
                g = self.named[n].grad
                g.copy_(flat[o:o + g.numel()].view_as(g) / world)
                o += g.numel()
        rest = [p.grad for n, p in self.named.items() if n not in self.seen and p.grad is not None]
        if rest:
            flat = torch.cat([g.reshape(-1) for g in rest])
            self.dist.all_reduce(flat, group=self.group)
            o = 0
            for g in rest:
                g.copy_(flat[o:o + g.numel()].view_as(g) / world)
                o += g.numel()
        self.pending, self.seen = [], set()


class MelLossFunction(torch.autograd.Function):
    """[lambda_l1 * l1, lambda_ssim * ssim] of add_mel_loss (tasks/tts/speech_base.py:219-257) and their gradient with respect to
    `mel_out`, both native (fse_mel_loss_forward / fse_mel_loss_backward, csrc/mel_loss.cu)."""

    @staticmethod
    def forward(ctx, mel_out, target, lam_l1: float, lam_ssim: float):
        _need_cuda(mel_out, target)
        mel_out, target = mel_out.contiguous().float(), target.contiguous().float()
        B, T, M = mel_out.shape
        nbytes = _lib.lib().fse_mel_loss_workspace_bytes(B, T, M)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=mel_out.device)
        losses = torch.empty(2, dtype=torch.float32, device=mel_out.device)
        want_grad = int(ctx.needs_input_grad[0])
        check(_lib.lib().fse_mel_loss_forward(_ptr(mel_out), _ptr(target), lam_l1, lam_ssim, _ptr(losses), want_grad, B, T, M, _ptr(ws), nbytes, _stream()))
        ctx.save_for_backward(mel_out, target, ws)
        ctx.lams = (lam_l1, lam_ssim)
        return losses

    @staticmethod
    def backward(ctx, dlosses):
        mel_out, target, ws = ctx.saved_tensors
        B, T, M = mel_out.shape
        grad = torch.empty_like(mel_out)
        dlosses = dlosses.contiguous().float()
        check(_lib.lib().fse_mel_loss_backward(_ptr(mel_out), _ptr(target), _ptr(dlosses), ctx.lams[0], ctx.lams[1], _ptr(grad), B, T, M, _ptr(ws),
                                                ws.numel(), _stream()))
        return grad, None, None, None


def mel_losses(mel_out: torch.Tensor, target: torch.Tensor, lambdas=(("l1", 0.5), ("ssim", 0.5))) -> Dict[str, torch.Tensor]:
    """add_mel_loss with the shipped `mel_losses: l1:0.5|ssim:0.5` (tasks/tts/speech_base.py:219-257, utils/metrics/ssim.py:24-44):
    one native forward for both terms (a term with weight 0 / not listed is dropped from the dict, as in the reference's loop)."""
    lam = dict(lambdas)
    unknown = set(lam) - {"l1", "ssim"}
    if unknown:
        raise NotImplementedError(f"mel loss terms {sorted(unknown)}: only l1 and ssim are on the B200 path")
    both = MelLossFunction.apply(mel_out, target, float(lam.get("l1", 0.0)), float(lam.get("ssim", 0.0)))
    return {name: both[i] for i, name in enumerate(("l1", "ssim")) if name in lam}


def train_step(denoise_fn, schedule: Dict[str, torch.Tensor], batch: Dict[str, torch.Tensor], optimizer, t: Optional[torch.Tensor] = None,
               noise: Optional[torch.Tensor] = None, reducer: Optional[BucketedAllReduce] = None) -> Dict[str, float]:
    """The denoiser branch of SpeechDenoiserTask.run_model(infer=False) (tasks/speech_editing/spec_denoiser.py:39-62 with
    GaussianDiffusion.forward(infer=False), spec_denoiser.py:168-176) for a given condition `cond` [B,T,H]:
      t ~ U{0..S};  x_t = q_sample(ref, t) * nonpadding;  x0 = denoise_fn(x_t, t, cond) * nonpadding;
      loss = l1 + ssim on the masked region;  backward;  (all-reduce);  optimizer step."""
    ref, mask, cond = batch["ref_mels"], batch["time_mel_masks"], batch["cond"]
    B, T, M = ref.shape
    nonpad = batch.get("nonpadding")
    nonpad = torch.ones(B, 1, 1, T, device=ref.device) if nonpad is None else nonpad.reshape(B, 1, 1, T).float()
    S = schedule["sqrt_alphas_cumprod"].numel() - 1
    if t is None:
        t = torch.randint(0, S + 1, (B,), device=ref.device)
    x_start = ref.transpose(1, 2)[:, None]
    if noise is None:
        noise = torch.randn_like(x_start)
    a = schedule["sqrt_alphas_cumprod"][t].view(B, 1, 1, 1)
    b = schedule["sqrt_one_minus_alphas_cumprod"][t].view(B, 1, 1, 1)
    x_t = (a * x_start + b * noise) * nonpad                                              # diffuse_fn / q_sample (:126-152)
    hook = reducer
    x0 = diffnet_train_forward(denoise_fn, x_t, t, cond.transpose(1, 2), grad_hook=hook) * nonpad
    mel_out = x0[:, 0].transpose(1, 2)
    m3 = mask.reshape(B, T, 1)
    losses = mel_losses(mel_out * m3, ref * m3)
    loss = sum(losses.values())
    optimizer.zero_grad(set_to_none=True)
    loss.backward()
    if reducer is not None:
        reducer.finish()
    optimizer.step()
    return {k: float(v.detach()) for k, v in losses.items()} | {"total": float(loss.detach())}
