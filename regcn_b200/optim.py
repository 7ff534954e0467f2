"""The optimisation step of the reference's training loop (src/main.py:194, 243-246):

    optimizer = torch.optim.Adam(model.parameters(), lr=args.lr, weight_decay=1e-5)
    ...
    loss.backward()
    torch.nn.utils.clip_grad_norm_(model.parameters(), args.grad_norm)
    optimizer.step()
    optimizer.zero_grad()

as two kernels over ONE flat fp32 buffer: a fixed-order gradient norm (`regcn_grad_norm`) and the Adam update with
the clip coefficient and the L2 weight decay folded in (`regcn_adam_step`).  Same call sequence:

    optimizer = regcn_b200.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-5)
    loss.backward(); regcn_b200.optim.clip_grad_norm_(optimizer, 1.0); optimizer.step(); optimizer.zero_grad()

Parameters and their gradients are re-pointed at views of the flat buffers on the first step (values preserved), so
autograd accumulates straight into the buffer the kernel reads.  Like torch.optim.Adam, parameters that received no
gradient are not updated (not even by weight decay).
"""
import torch

from . import _lib
from ._lib import call, ptr

F32 = torch.float32


class Adam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_grad_norm=None):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.max_grad_norm = max_grad_norm
        self._pending_clip = None
        self.step_count = 0
        self._flat = None
        self.total_norm = None          # device float: the gradient norm of the last clipped step

    # ---- flat buffers -------------------------------------------------------------------------------------------
    def _build(self):
        act = [p for p in self.params if p.grad is not None]
        if not act:
            raise RuntimeError("regcn_b200.optim.Adam.step(): no parameter has a gradient")
        dev = act[0].device
        if not act[0].is_cuda:
            raise RuntimeError("regcn_b200: kernels take CUDA tensors only (no CPU fallback)")
        offs, tot = [], 0
        for p in act:
            if p.dtype != F32 or p.device != dev:
                raise TypeError("regcn_b200.optim.Adam: parameters must be float32 on one device")
            offs.append(tot)
            tot += (p.numel() + 3) // 4 * 4                      # 16-byte aligned views
        flat_p = torch.zeros(tot, device=dev, dtype=F32)
        flat_g = torch.zeros(tot, device=dev, dtype=F32)
        views_g = []
        with torch.no_grad():
            for p, o in zip(act, offs):
                vp = flat_p[o:o + p.numel()].view(p.shape)
                vg = flat_g[o:o + p.numel()].view(p.shape)
                vp.copy_(p.data)
                vg.copy_(p.grad)
                p.data = vp
                p.grad = vg
                views_g.append(vg)
        nb = _lib.load().regcn_adam_workspace_bytes()
        self._flat = dict(p=flat_p, g=flat_g, m=torch.zeros_like(flat_p), v=torch.zeros_like(flat_p), act=act,
                          views_g=views_g, ids={id(p) for p in act}, n=tot, ws=torch.empty((nb + 7) // 8, device=dev, dtype=torch.float64),
                          ws_bytes=nb)
        self.total_norm = torch.zeros(1, device=dev, dtype=F32)

    def _sync_grads(self):
        """Gradients must live in the flat buffer; re-attach the views if something replaced or dropped p.grad."""
        f = self._flat
        late = [p for p in self.params if p.grad is not None and id(p) not in f["ids"]]
        if late:
            # torch.optim.Adam would start these parameters' moments and step count now; the fused kernel keeps ONE step
            # count for the flat buffer, so silently skipping them (or sharing the count) would diverge from the reference
            raise RuntimeError(f"regcn_b200.optim.Adam: {len(late)} parameter(s) received their first gradient after the first "
                               "step() (e.g. a loss head switched on later); build the optimizer after the first backward of "
                               "the full loss, or create a new optimizer")
        with torch.no_grad():
            for p, vg in zip(f["act"], f["views_g"]):
                if p.grad is None:
                    vg.zero_()
                    p.grad = vg
                elif p.grad.data_ptr() != vg.data_ptr():
                    vg.copy_(p.grad)
                    p.grad = vg

    # ---- torch.optim surface ------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none=False):
        if self._flat is None:
            for p in self.params:
                p.grad = None
            return
        self._flat["g"].zero_()

    def clip_grad_norm_(self, max_norm):
        """Records the clip; the norm and the scaling run inside the next step() (no extra pass over the gradients).
        Returns the device tensor that will hold the total norm after step()."""
        self._pending_clip = float(max_norm)
        return self.total_norm

    @torch.no_grad()
    def step(self):
        if self._flat is None:
            self._build()
        else:
            self._sync_grads()
        f = self._flat
        clip = self._pending_clip if self._pending_clip is not None else self.max_grad_norm
        self._pending_clip = None
        self.step_count += 1
        tn = None
        if clip is not None and clip > 0:
            call("regcn_grad_norm", ptr(f["g"]), f["n"], ptr(self.total_norm), ptr(f["ws"]), f["ws_bytes"])
            tn = self.total_norm
        call("regcn_adam_step", ptr(f["p"]), ptr(f["g"]), ptr(f["m"]), ptr(f["v"]), f["n"], self.lr, self.betas[0],
             self.betas[1], self.eps, self.weight_decay, self.step_count, float(clip) if tn is not None else 0.0, ptr(tn))
        for p in f["act"]:
            torch.autograd.graph.increment_version(p)     # derived operand caches key on the version counter


def clip_grad_norm_(optimizer, max_norm):
    """torch.nn.utils.clip_grad_norm_ of the reference's loop (src/main.py:244), fused into optimizer.step()."""
    return optimizer.clip_grad_norm_(max_norm)
