"""MNIST MVAE on the B200-native library: module surface of the reference + the fused train step.

Reference surface kept (mnist/model.py, mnist/train.py):
    MVAE / MultimodalVAE(n_latents).forward(image=None, text=None) -> (recon_image, recon_text, mu, logvar)
    state_dict keys identical to the reference's (image_encoder.net.0.weight, ...)
The hot path is `MVAETrainer.step(image, text)`: ONE C-ABI call (mvae_mnist_step) that enqueues the whole
three-term ELBO step - forward, backward, Adam - as hand-written sm_100a kernels, optionally replayed as a
CUDA graph.  There is no PyTorch fallback: without the CUDA library every call raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib

TERMS = {"joint": _lib.TERM_JOINT, "image": _lib.TERM_IMAGE, "text": _lib.TERM_TEXT}
# "tf32x3": fp32 storage with error-compensated 3xTF32 GEMMs (hi/lo operand split) - fp32-grade parity mode
_DTYPES = {"tf32": _lib.DT_F32, "fp32": _lib.DT_F32, "bf16": _lib.DT_BF16, "tf32x3": _lib.DT_F32X3}


def _stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def tensor_table(n_latents: int):
    """[(name, kind, shape, offset)] in the reference's state_dict order, from the C library."""
    lib = _lib.load()
    out = []
    for i in range(lib.mvae_mnist_num_tensors()):
        ti = _lib.TensorInfo()
        _lib.check(lib.mvae_mnist_tensor_info(n_latents, i, C.byref(ti)), "mvae_mnist_tensor_info")
        shape = tuple(int(ti.shape[d]) for d in range(ti.ndim))
        out.append((ti.name.decode(), int(ti.kind), shape, int(ti.offset)))
    return out


def sizes(n_latents: int, batch: int, dtype: int) -> _lib.MnistSizeInfo:
    si = _lib.MnistSizeInfo()
    _lib.check(_lib.load().mvae_mnist_sizes(n_latents, batch, dtype, C.byref(si)), "mvae_mnist_sizes")
    return si


class _Leaf(nn.Module):
    """Parameter/buffer holder standing in for nn.Linear / nn.BatchNorm1d / nn.Embedding (keys only)."""


class _Net(nn.Module):
    pass


class _Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.net = _Net()


class MVAE(nn.Module):
    """Drop-in for the reference's MultimodalVAE (mnist/model.py:14-96).

    All parameters are views into one flat fp32 device buffer (`flat_params`), gradients into `flat_grads`.
    `precision`: "tf32" (fp32 storage, tensor cores in tf32), "tf32x3" (fp32 storage, error-compensated
    3xTF32 GEMMs: the parity mode that meets rtol 1e-3 on every gradient) or "bf16" (the fast path).
    """

    def __init__(self, n_latents: int = 20, precision: str = "tf32", device: Optional[torch.device] = None,
                 poe_mode: str = "ref", prior_expert: bool = False, seed: int = 0):
        super().__init__()
        if precision not in _DTYPES:
            raise ValueError("precision must be one of %s" % sorted(_DTYPES))
        self.n_latents = int(n_latents)
        self.precision = precision
        self.dtype_code = _DTYPES[precision]
        self.poe_mode = {"ref": _lib.POE_REF, "precision": _lib.POE_PRECISION}[poe_mode]
        self.prior_expert = bool(prior_expert)
        self.noise_seed = int(seed)
        need = 8 if self.dtype_code == _lib.DT_BF16 else 4
        if self.n_latents % need:
            raise ValueError("n_latents must be a multiple of %d for precision=%s (TMA 16-byte rows)" % (need, precision))
        dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        if dev.type != "cuda":
            raise RuntimeError("mvae_b200 has no CPU path: a CUDA (sm_100) device is required")
        _lib.check(_lib.load().mvae_device_check(dev.index or 0), "mvae_device_check")
        self.device_ = dev
        si = sizes(self.n_latents, 2, self.dtype_code)
        self.flat_params = torch.zeros(si.param_floats, device=dev, dtype=torch.float32)
        self.flat_grads = torch.zeros(si.param_floats, device=dev, dtype=torch.float32)
        self.flat_buffers = torch.zeros(si.buffer_floats, device=dev, dtype=torch.float32)
        self.flat_nbt = torch.zeros(si.num_bn, device=dev, dtype=torch.int64)
        self.flat_params_bf16 = (torch.zeros(si.param_floats, device=dev, dtype=torch.bfloat16)
                                 if self.dtype_code == _lib.DT_BF16 else None)
        self._table = tensor_table(self.n_latents)
        self.image_encoder, self.image_decoder = _Block(), _Block()
        self.text_encoder, self.text_decoder = _Block(), _Block()
        from .functional import ProductOfExperts
        self.experts = ProductOfExperts("ref" if self.poe_mode == _lib.POE_REF else "precision", self.prior_expert)
        for name, kind, shape, off in self._table:
            block, _, idx, leaf = name.split(".")
            net = getattr(self, block).net
            if not hasattr(net, idx):
                net.add_module(idx, _Leaf())
            holder = getattr(net, idx)
            numel = 1
            for s in shape:
                numel *= s
            if kind == 0:
                p = nn.Parameter(self.flat_params[off:off + numel].view(shape))
                p.grad = self.flat_grads[off:off + numel].view(shape)
                holder.register_parameter(leaf, p)
            elif kind == 1:
                holder.register_buffer(leaf, self.flat_buffers[off:off + numel].view(shape))
            else:
                holder.register_buffer(leaf, self.flat_nbt[off])
        self.reset_parameters()
        self._ws: Dict[Tuple[int, int], torch.Tensor] = {}
        self._injected_noise = []
        # two device clocks (so that CUDA graphs stay valid): [0] the Philox / noise counter, ticked by every forward-type
        # call; [1] Adam's bias-correction step, ticked only by optimizer steps (checkpointed as "step")
        self._counters = torch.zeros(2, device=dev, dtype=torch.int32)
        self._step_counter = self._counters[0:1]
        self._adam_counter = self._counters[1:2]

    # ------------------------------------------------------------------ parameters
    @torch.no_grad()
    def reset_parameters(self, seed: Optional[int] = None) -> None:
        """PyTorch default initialisers of nn.Linear / nn.Embedding / nn.BatchNorm1d."""
        g = torch.Generator().manual_seed(1234 if seed is None else seed)
        sd = self.state_dict()
        for name, kind, shape, _ in self._table:
            t = sd[name]
            if kind == 2:
                t.zero_()
            elif name.endswith("running_mean"):
                t.zero_()
            elif name.endswith("running_var"):
                t.fill_(1.0)
            elif ".net.1." in name or ".net.4." in name:
                t.fill_(1.0 if name.endswith("weight") else 0.0)
            elif name == "text_encoder.net.0.weight":
                t.copy_(torch.randn(shape, generator=g))
            else:
                fan_in = shape[1] if len(shape) == 2 else sd[name[:-4] + "weight"].shape[1]
                bound = 1.0 / fan_in ** 0.5
                t.copy_((torch.rand(shape, generator=g) * 2 - 1) * bound)
        self.sync_low_precision()

    def sync_low_precision(self) -> None:
        """Refresh the bf16 mirror of the parameters (after load_state_dict or any manual edit)."""
        if self.flat_params_bf16 is not None:
            _lib.check(_lib.load().mvae_cast_f32_to_bf16(
                C.c_void_p(self.flat_params.data_ptr()), C.c_void_p(self.flat_params_bf16.data_ptr()),
                C.c_int64(self.flat_params.numel()), _stream_ptr()), "mvae_cast_f32_to_bf16")

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        out = super().load_state_dict(state_dict, strict=strict, assign=False)
        self.sync_low_precision()
        return out

    # ------------------------------------------------------------------ workspace / conversions
    def workspace(self, batch: int) -> torch.Tensor:
        key = (batch, self.dtype_code)
        if key not in self._ws:
            si = sizes(self.n_latents, batch, self.dtype_code)
            self._ws[key] = torch.empty(si.workspace_bytes + 256, device=self.device_, dtype=torch.uint8)
        ws = self._ws[key]
        off = (-ws.data_ptr()) % 256
        return ws[off:]

    def debug_buffer(self, name: str, batch: int, shape, dtype=None) -> torch.Tensor:
        """View of a named intermediate buffer of the last step (tests / bring-up)."""
        off = _lib.load().mvae_mnist_workspace_offset(name.encode(), batch, self.n_latents, self.dtype_code)
        if off < 0:
            raise KeyError(name)
        dtype = dtype or self.act_dtype()
        numel = 1
        for s_ in shape:
            numel *= s_
        nbytes = numel * torch.empty((), dtype=dtype).element_size()
        return self.workspace(batch)[off:off + nbytes].view(dtype).view(shape)

    def act_dtype(self) -> torch.dtype:
        return torch.bfloat16 if self.dtype_code == _lib.DT_BF16 else torch.float32

    def to_act(self, image: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """image.view(-1, 784) in the storage dtype of the tensor-core path (mnist/train.py:131).  `out`: uint8 pixels are
        converted straight into this buffer (the graph trainer's static input)."""
        x = image.reshape(-1, 784)
        if x.dtype == torch.uint8:
            if out is None:
                out = torch.empty(x.shape, device=self.device_, dtype=self.act_dtype())
            xd = x.to(self.device_, non_blocking=True).contiguous()
            f32 = out.data_ptr() if self.dtype_code != _lib.DT_BF16 else None
            b16 = out.data_ptr() if self.dtype_code == _lib.DT_BF16 else None
            _lib.check(_lib.load().mvae_u8_to_act(C.c_void_p(xd.data_ptr()), C.c_void_p(f32), C.c_void_p(b16),
                                                  C.c_int64(xd.numel()), C.c_float(1.0 / 255.0), _stream_ptr()),
                       "mvae_u8_to_act")
            return out
        x = x.to(self.device_, non_blocking=True)
        if x.dtype == self.act_dtype():
            return x.contiguous()
        if x.dtype == torch.float32 and self.dtype_code == _lib.DT_BF16:
            x = x.contiguous()
            out = torch.empty(x.shape, device=self.device_, dtype=torch.bfloat16)
            _lib.check(_lib.load().mvae_cast_f32_to_bf16(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()),
                                                         C.c_int64(x.numel()), _stream_ptr()), "mvae_cast_f32_to_bf16")
            return out
        raise TypeError("image dtype %s not supported for precision %s" % (x.dtype, self.precision))

    # ------------------------------------------------------------------ the C call
    def _build_args(self, image, text, terms: Sequence[int], lambdas, kl_weights, eps=None, backward=False,
                    zero_grad=True, adam=None, outputs=False, grad_scale=1.0, losses=None, workspace=None,
                    grads=None, extra=None):
        B = image.shape[0]
        G = len(terms)
        if image.dtype != self.act_dtype() or not image.is_contiguous() or image.shape[1] != 784:
            raise ValueError("image must be a contiguous [B, 784] %s tensor (use MVAE.to_act)" % self.act_dtype())
        if text.dtype != torch.int64 or text.shape[0] != B:
            raise ValueError("text must be an int64 [B] tensor")
        a = _lib.MnistStepArgs()
        a.batch, a.n_latents, a.dtype, a.n_terms = B, self.n_latents, self.dtype_code, G
        for g in range(G):
            a.term_type[g] = terms[g]
            a.lambda_image[g], a.lambda_text[g] = float(lambdas[g][0]), float(lambdas[g][1])
            a.kl_weight[g] = float(kl_weights[g])
        a.poe_mode, a.prior_expert, a.poe_eps = self.poe_mode, int(self.prior_expert), 1e-8
        a.image, a.text = image.data_ptr(), text.data_ptr()
        a.eps = None if eps is None else eps.data_ptr()
        a.seed = self.noise_seed
        a.params = self.flat_params.data_ptr()
        a.params_bf16 = None if self.flat_params_bf16 is None else self.flat_params_bf16.data_ptr()
        a.buffers = self.flat_buffers.data_ptr()
        a.num_batches_tracked = self.flat_nbt.data_ptr()
        a.grads = (self.flat_grads if grads is None else grads).data_ptr()
        a.do_backward, a.zero_grad = int(backward), int(zero_grad)
        a.adam_step = self._adam_counter.data_ptr()
        a.noise_step = self._step_counter.data_ptr()
        a.grad_scale = float(grad_scale)
        if adam is not None:
            a.do_adam = 1
            a.adam_m, a.adam_v = adam["m"].data_ptr(), adam["v"].data_ptr()
            a.lr, a.beta1, a.beta2, a.adam_eps = adam["lr"], adam["betas"][0], adam["betas"][1], adam["eps"]
        ws = self.workspace(B) if workspace is None else workspace
        a.workspace, a.workspace_bytes = ws.data_ptr(), ws.numel()
        if losses is None:
            losses = torch.empty(G, 4, device=self.device_, dtype=torch.float32)
        a.out_losses = losses.data_ptr()
        outs = None
        if outputs:
            ri = torch.empty(G * B, 784, device=self.device_, dtype=self.act_dtype())
            rt = torch.empty(G * B, 10, device=self.device_, dtype=torch.float32)
            mu = torch.empty(G, B, self.n_latents, device=self.device_, dtype=torch.float32)
            lv = torch.empty_like(mu)
            a.out_recon_image, a.out_recon_text = ri.data_ptr(), rt.data_ptr()
            a.out_mu, a.out_logvar = mu.data_ptr(), lv.data_ptr()
            outs = (ri, rt, mu, lv)
        for k, v in (extra or {}).items():
            setattr(a, k, v)
        return a, losses, outs

    def _run(self, image, text, terms, lambdas, kl_weights, **kw):
        a, losses, outs = self._build_args(image, text, terms, lambdas, kl_weights, **kw)
        _lib.check(_lib.load().mvae_mnist_step(C.byref(a), _stream_ptr()), "mvae_mnist_step")
        return losses, outs

    def profile(self, image, text, terms, lambdas, kl_weights, **kw):
        """One step with a CUDA-event pair around every kernel launch: [(label, milliseconds)]."""
        a, _, _ = self._build_args(image, text, terms, lambdas, kl_weights, **kw)
        n_max, stride = 96, 64
        labels = C.create_string_buffer(n_max * stride)
        ms = (C.c_float * n_max)()
        n = C.c_int(0)
        _lib.check(_lib.load().mvae_mnist_step_profile(C.byref(a), _stream_ptr(), n_max, labels, stride, ms,
                                                       C.byref(n)), "mvae_mnist_step_profile")
        raw = labels.raw
        return [(raw[i * stride:(i + 1) * stride].split(b"\0", 1)[0].decode(), float(ms[i])) for i in range(n.value)]

    # ------------------------------------------------------------------ reference surface
    def new_workspace(self, batch: int) -> torch.Tensor:
        si = sizes(self.n_latents, batch, self.dtype_code)
        ws = torch.empty(si.workspace_bytes + 256, device=self.device_, dtype=torch.uint8)
        return ws[(-ws.data_ptr()) % 256:]

    def _prep_inputs(self, image, text):
        assert image is not None or text is not None
        term = TERMS["joint"] if (image is not None and text is not None) else (
            TERMS["image"] if image is not None else TERMS["text"])
        B = image.shape[0] if image is not None else text.shape[0]
        x = self.to_act(image) if image is not None else torch.zeros(B, 784, device=self.device_, dtype=self.act_dtype())
        y = (text.to(self.device_).long().contiguous() if text is not None
             else torch.zeros(B, device=self.device_, dtype=torch.int64))
        return term, B, x, y

    def forward(self, image=None, text=None):
        """MultimodalVAE.forward (mnist/model.py:53-84): (recon_image probs, recon_text log-probs, mu, logvar).

        train(): batch statistics (running statistics advance), z sampled; the outputs carry an autograd graph
        whose backward runs the library's backward kernels, so the reference loop (three forwards,
        `loss_function`, `.backward()`, any torch optimizer) works unchanged.  eval(): running statistics, z = mu.
        The fast path for training is MVAETrainer.step, which fuses the three terms."""
        term, B, x, y = self._prep_inputs(image, text)
        self.sync_low_precision()   # the parameters may have been stepped by a torch optimizer (bf16 GEMMs read the mirror)
        if not self.training:
            with torch.no_grad():
                _, outs = self._run(x, y, [term], [(0.0, 0.0)], [0.0], outputs=True, extra={"eval_mode": 1})
            ri, rt, mu, lv = outs
            return ri, rt, mu[0], lv[0]
        if self._injected_noise:   # tests: the N(0,1) draw of reparametrize (mnist/model.py:27) can be injected
            eps = self._injected_noise.pop(0).to(self.device_, torch.float32).reshape(1, B, self.n_latents).contiguous()
        else:
            eps = torch.randn(1, B, self.n_latents, device=self.device_, dtype=torch.float32)
        if torch.is_grad_enabled():
            return _MVAEForwardFn.apply(self, x, y, term, eps, *self.parameters())
        _, outs = self._run(x, y, [term], [(0.0, 0.0)], [0.0], eps=eps, outputs=True)
        ri, rt, mu, lv = outs
        return ri, rt, mu[0], lv[0]

    # sub-calls used by the reference's sampling / visualisation scripts (mnist/sample.py:78-116, manifold.py:49)
    @torch.no_grad()
    def encode_image(self, x):
        """ImageEncoder (mnist/model.py:99-117): (mu, logvar) of the image expert."""
        _, B, xa, y = self._prep_inputs(x, None)
        self.sync_low_precision()
        self._run(xa, y, [TERMS["image"]], [(0.0, 0.0)], [0.0], extra={"eval_mode": int(not self.training)})
        enc = self.debug_buffer("enc", B, (B, 2 * self.n_latents), torch.float32).clone()
        return enc[:, :self.n_latents], enc[:, self.n_latents:]

    @torch.no_grad()
    def encode_text(self, x):
        """TextEncoder (mnist/model.py:138-153): (mu, logvar) of the text expert."""
        _, B, xa, y = self._prep_inputs(None, x)
        self.sync_low_precision()
        self._run(xa, y, [TERMS["text"]], [(0.0, 0.0)], [0.0], extra={"eval_mode": int(not self.training)})
        table = self.debug_buffer("txt_table", B, (10, 2 * self.n_latents), torch.float32).clone()
        out = table[y]
        return out[:, :self.n_latents], out[:, self.n_latents:]

    @torch.no_grad()
    def _decode(self, z):
        z = z.to(self.device_, torch.float32).contiguous()
        B = z.shape[0]
        x = torch.zeros(B, 784, device=self.device_, dtype=self.act_dtype())
        y = torch.zeros(B, device=self.device_, dtype=torch.int64)
        self.sync_low_precision()
        _, outs = self._run(x, y, [TERMS["joint"]], [(0.0, 0.0)], [0.0], outputs=True,
                            extra={"eval_mode": int(not self.training), "z_in": z.data_ptr()})
        return outs[0], outs[1]

    @torch.no_grad()
    def decode_losses(self, z, image, text):
        """Reconstruction losses of latents z against (image, text) in ONE fused launch sequence (decoders + the BCE
        epilogue + the NLL kernel, no probabilities materialised): returns the device tensor [bce_mean, nll_mean] with
        the means of F.binary_cross_entropy over B*784 pixels and F.nll_loss over B labels - the inner loop of
        mnist/loglikelihood.py:46-52."""
        _, B, x, y = self._prep_inputs(image, text)
        z = z.to(self.device_, torch.float32).contiguous()
        self._keep_z = z
        self.sync_low_precision()
        losses, _ = self._run(x, y, [TERMS["joint"]], [(1.0, 1.0)], [0.0],
                              extra={"eval_mode": int(not self.training), "z_in": z.data_ptr()})
        return losses[0, 1:3]

    def decode_image(self, x):
        """ImageDecoder (mnist/model.py:120-135): pixel probabilities for latents x."""
        return self._decode(x)[0]

    def decode_text(self, x):
        """TextDecoder (mnist/model.py:156-170): label log-probabilities for latents x."""
        return self._decode(x)[1]

    def reparametrize(self, mu, logvar):
        """mnist/model.py:24-30 (host-side convenience, not on the hot path)."""
        if self.training:
            return torch.randn_like(mu) * torch.exp(0.5 * logvar) + mu
        return mu

    def prior(self, size, use_cuda=True):
        """mnist/model.py:44-51."""
        return torch.zeros(size, device=self.device_), torch.zeros(size, device=self.device_)

    def gen_latents(self, image, text):
        """mnist/model.py:86-96: a sample from the joint posterior q(z | image, text)."""
        with torch.no_grad():
            mu_i, lv_i = self.encode_image(image)
            mu_t, lv_t = self.encode_text(text)
            mu, lv = self.experts(torch.stack((mu_i, mu_t)), torch.stack((lv_i, lv_t)))
            return self.reparametrize(mu, lv)


class _MVAEForwardFn(torch.autograd.Function):
    """Autograd bridge of MVAE.forward: forward = one forward-only C call on a private workspace; backward = one
    backward-only C call (phase 2) fed with the upstream gradients of (recon_image, recon_text, mu, logvar)."""

    @staticmethod
    def forward(ctx, model, x, y, term, eps, *params):
        B = x.shape[0]
        ws = model.new_workspace(B)
        _, outs = model._run(x, y, [term], [(0.0, 0.0)], [0.0], eps=eps, outputs=True, workspace=ws)
        ri, rt, mu, lv = outs
        ctx.model, ctx.ws, ctx.x, ctx.y, ctx.term, ctx.eps = model, ws, x, y, term, eps
        ctx.save_for_backward(ri)
        return ri, rt, mu[0].clone(), lv[0].clone()

    @staticmethod
    def backward(ctx, g_ri, g_rt, g_mu, g_lv):
        m = ctx.model
        (ri,) = ctx.saved_tensors

        def c(t, dtype):
            return None if t is None else t.to(dtype).contiguous()

        g_ri, g_rt = c(g_ri, m.act_dtype()), c(g_rt, torch.float32)
        g_mu, g_lv = c(g_mu, torch.float32), c(g_lv, torch.float32)
        grads = torch.zeros_like(m.flat_params)
        extra = {"phase": 2, "d_recon_image": _lib.ptr(g_ri), "d_recon_text": _lib.ptr(g_rt),
                 "d_mu": _lib.ptr(g_mu), "d_logvar": _lib.ptr(g_lv), "out_recon_image": ri.data_ptr()}
        a, _, _ = m._build_args(ctx.x, ctx.y, [ctx.term], [(0.0, 0.0)], [0.0], eps=ctx.eps, backward=True,
                                zero_grad=False, workspace=ctx.ws, grads=grads, extra=extra)
        _lib.check(_lib.load().mvae_mnist_step(C.byref(a), _stream_ptr()), "mvae_mnist_step(backward)")
        out = []
        for name, kind, shape, off in m._table:
            if kind == 0:
                numel = 1
                for s_ in shape:
                    numel *= s_
                out.append(grads[off:off + numel].view(shape))
        return (None, None, None, None, None, *out)


MultimodalVAE = MVAE  # the reference's class name (mnist/model.py:14)


class MVAETrainer:
    """The fused three-term ELBO step of mnist/train.py:127-153 (and the weak-supervision variants
    mnist/modal_weak.py:69-100) as one launch sequence, with Adam state living next to the flat parameters."""

    def __init__(self, model: MVAE, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 use_cuda_graph: bool = False, grad_scale: float = 1.0):
        self.model = model
        self.adam = {"m": torch.zeros_like(model.flat_params), "v": torch.zeros_like(model.flat_params),
                     "lr": lr, "betas": betas, "eps": eps}
        self.use_cuda_graph = use_cuda_graph
        self.grad_scale = grad_scale
        self._graphs = {}
        self.last_graph_launches = 0
        # graph replay: two sets of static input buffers (and one captured graph per set) alternate, so that the next
        # step's inputs are staged on a copy stream while the current step still reads its own
        self._slot = 0
        self._stage_stream = None

    def _graph_input(self, image):
        """What a graph step takes as its image input: a uint8 device batch goes in as it is (_stage converts it straight
        into the graph's static activation buffer, beside the previous step), anything else through to_act()."""
        if image.dtype == torch.uint8 and image.is_cuda:
            return image.reshape(-1, 784).contiguous()
        return self.model.to_act(image)

    def _stage(self, ent, x, y, eps, ready=None):
        """Copy the step's inputs into the entry's static buffers on the staging stream: behind the last replay that
        used this entry (two steps ago), i.e. beside the previous step - provided the caller says when the inputs are
        complete: `ready` = a torch.cuda.Event recorded by their producer, True (complete since long, e.g. a resident
        pool), or None: ordered behind everything enqueued on the current stream so far (always safe, no overlap)."""
        dev = self.model.device_
        if self._stage_stream is None:
            self._stage_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        st = self._stage_stream
        if ent.get("fresh") or ready is None:
            st.wait_stream(main)
        else:
            st.wait_event(ent["free"])
            if ready is not True:
                st.wait_event(ready)
        with torch.cuda.stream(st):
            if x.dtype == torch.uint8 and ent["x"].dtype != torch.uint8:
                self.model.to_act(x, out=ent["x"])   # mvae_u8_to_act on the staging stream
            else:
                ent["x"].copy_(x, non_blocking=True)
            ent["y"].copy_(y, non_blocking=True)
            if eps is not None:
                ent["eps"].copy_(eps, non_blocking=True)
            ent["ready"].record(st)
        for t in (x, y, eps):
            if t is not None:
                t.record_stream(st)
        main.wait_event(ent["ready"])
        ent["fresh"] = False

    @staticmethod
    def mnist_kl_weight(batch: int, annealing_factor: float = 1.0) -> float:
        """KLD /= batch_size * (784 / 3)   (mnist/train.py:80)."""
        return annealing_factor * 3.0 / (784.0 * batch)

    def _norm(self, terms, batch, annealing_factor):
        tt = [TERMS[t] if isinstance(t, str) else int(t) for t in terms]
        return tt, [self.mnist_kl_weight(batch, annealing_factor)] * len(tt)

    def step(self, image, text, eps=None, terms=("joint", "image", "text"), lambdas=((1.0, 1.0),) * 3,
             annealing_factor: float = 1.0, update: bool = True, outputs: bool = False, zero_grad: bool = True, ready=None):
        """One training step.  Returns (losses [n_terms, 4] device tensor: total / image / text / KL per term,
        outputs or None).  `image`: [B,784] (or [B,1,28,28]) uint8 / float32 / storage dtype, host or device.
        `ready` (graph replay only): when the device inputs are complete - see _stage()."""
        m = self.model
        y = text.to(m.device_, non_blocking=True).long().contiguous()
        if eps is not None:
            eps = eps.to(m.device_, torch.float32).contiguous()
        if self.use_cuda_graph and not outputs:
            return self._graph_step(self._graph_input(image), y, eps, tuple(terms), tuple(map(tuple, lambdas)),
                                    annealing_factor, update, zero_grad, ready=ready), None
        x = m.to_act(image)
        tt, klw = self._norm(terms, x.shape[0], annealing_factor)
        return m._run(x, y, tt, lambdas, klw, eps=eps, backward=True, zero_grad=zero_grad,
                      adam=self.adam if update else None, outputs=outputs, grad_scale=self.grad_scale)

    #: (row selector, terms, lambdas) of the three presence classes of step_masked
    _MASK_CLASSES = (("paired", ("joint", "image", "text"), ((1.0, 1.0), (1.0, 1.0), (0.0, 1.0))),
                     ("image_only", ("image",), ((1.0, 0.0),)),
                     ("text_only", ("text",), ((0.0, 1.0),)))

    def step_masked(self, image, text, has_image, has_text, eps=None, annealing_factor: float = 1.0,
                    update: bool = True):
        """One optimizer step over a batch that MIXES paired and unpaired samples (the per-sample
        missing-modality mask end to end; the reference only flips a coin per batch, mnist/paired_weak.py:82-104).

        `has_image`, `has_text`: [B] bool.  Rows are compacted by presence class on the device and each class runs
        the fused step on its own rows: paired rows -> joint + image + text terms with the lambdas of
        mnist/paired_weak.py:85-93, image-only rows -> the image term (:97-99), text-only rows -> the text term
        (:100-102); rows with neither modality are skipped.  Every loss term is a mean over its own rows and
        BatchNorm only ever sees present rows - exactly what the reference computes when it is fed the three
        subsets as three batches - but the gradients accumulate and ONE Adam update follows.
        `eps`: optional [3, B, n] noise per (term, row).  Returns {class: losses [n_terms, 4]} (device tensors)."""
        m = self.model
        hi = torch.as_tensor(has_image).to(m.device_).bool().reshape(-1)
        ht = torch.as_tensor(has_text).to(m.device_).bool().reshape(-1)
        x = m.to_act(image)
        y = text.to(m.device_, non_blocking=True).long().contiguous()
        if hi.numel() != x.shape[0] or ht.numel() != x.shape[0]:
            raise ValueError("has_image / has_text must have one entry per sample")
        if eps is not None:
            eps = eps.to(m.device_, torch.float32)
        rows = {"paired": hi & ht, "image_only": hi & ~ht, "text_only": ~hi & ht}
        term_index = {"joint": 0, "image": 1, "text": 2}
        out, first, ran = {}, True, 0
        ws = m.workspace(x.shape[0])  # the full batch's workspace is large enough for every subset
        for j, (name, terms, lambdas) in enumerate(self._MASK_CLASSES):
            idx = torch.nonzero(rows[name]).reshape(-1)  # data-dependent size: one host sync per class
            if idx.numel() == 0:
                continue
            e = None
            if eps is not None:
                e = torch.stack([eps[term_index[t]].index_select(0, idx) for t in terms]).contiguous()
            tt, klw = self._norm(terms, int(idx.numel()), annealing_factor)
            # a distinct Philox stream per class; Adam's clock ticks once, with the first class that runs
            losses, _ = m._run(x.index_select(0, idx), y.index_select(0, idx), tt, lambdas, klw, eps=e, backward=True,
                               zero_grad=first, adam=None, grad_scale=self.grad_scale, workspace=ws,
                               extra={"seed": (m.noise_seed + 0x9E3779B1 * (j + 1)) & 0x7FFFFFFFFFFFFFFF,
                                      "advance_adam_step": int(update and first)})
            out[name] = losses
            first = False
            ran += 1
        self._reduce_gradients()
        if update and ran > 0:
            from . import _ops
            a = self.adam
            _ops.adam_step(m.flat_params, m.flat_grads, a["m"], a["v"], m.flat_params_bf16, m.flat_params.numel(), a["lr"],
                           a["betas"][0], a["betas"][1], a["eps"], m._adam_counter, self.grad_scale, zero_grad=False)
        return out

    def _reduce_gradients(self) -> None:
        """Hook between the backward of an accumulated step and its Adam update (data parallel: the all-reduce)."""

    _MAX_GRAPHS = 32   # captured graphs kept per trainer (least recently used are dropped: each owns static buffers)

    def _graph_cache_get(self, cache, key):
        ent = cache.get(key)
        if ent is not None:
            cache[key] = cache.pop(key)   # move to the end: most recently used
        return ent

    def _graph_cache_put(self, cache, key, ent):
        cache[key] = ent
        while len(cache) > self._MAX_GRAPHS:
            cache.pop(next(iter(cache)))

    def _graph_step(self, x, y, eps, terms, lambdas, annealing_factor, update, zero_grad, ready=None):
        """Replay of the step as one CUDA graph (static input buffers; one graph per configuration)."""
        m = self.model
        # every scalar that capture bakes into kernel arguments is part of the key (Adam hyper-parameters included:
        # adjust_learning_rate / annealing schedules mutate them between steps); the cache is bounded (LRU)
        a_ = self.adam
        self._slot ^= 1
        key = (x.shape[0], terms, lambdas, float(annealing_factor), bool(update), bool(zero_grad), eps is not None,
               float(a_["lr"]), tuple(map(float, a_["betas"])), float(a_["eps"]), float(self.grad_scale), x.dtype, self._slot)
        ent = self._graph_cache_get(self._graphs, key)
        if ent is None:
            # uint8 pixels never enter the graph: the conversion to the storage dtype IS the staging copy (stage stream,
            # beside the previous step), so the step's first node is the encoder chain itself
            sx = torch.empty(x.shape, device=m.device_, dtype=m.act_dtype()) if x.dtype == torch.uint8 else torch.empty_like(x)
            sy = torch.empty_like(y)
            se = torch.empty_like(eps) if eps is not None else None
            losses = torch.empty(len(terms), 4, device=m.device_, dtype=torch.float32)
            tt, klw = self._norm(terms, x.shape[0], annealing_factor)
            kw = dict(eps=se, backward=True, zero_grad=zero_grad, adam=self.adam if update else None,
                      grad_scale=self.grad_scale, losses=losses)
            # capture does not execute: the caller-visible step happens at the first replay below
            m.workspace(x.shape[0])
            torch.cuda.synchronize()
            lib = _lib.load()
            before = lib.mvae_launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                m._run(sx, sy, tt, lambdas, klw, **kw)
            ent = {"graph": graph, "x": sx, "y": sy, "eps": se, "losses": losses, "fresh": True,
                   "free": torch.cuda.Event(), "ready": torch.cuda.Event(),
                   "launches": int(lib.mvae_launch_count() - before) + (1 if x.dtype == torch.uint8 else 0)}
            self._graph_cache_put(self._graphs, key, ent)
        self._stage(ent, x, y, eps, ready)
        ent["graph"].replay()
        ent["free"].record(torch.cuda.current_stream(m.device_))
        self.last_graph_launches = ent["launches"]
        return ent["losses"]


class HostPipeline:
    """Feeds host batches to a trainer with the copies overlapped with compute (the public end-to-end path).

    Batches are (image uint8/float32 [B,784] or [B,1,28,28], labels int64 [B]) in PINNED host memory.  Batch i+1
    is uploaded on a copy stream while step i runs; each step's [n_terms, 4] loss tensor is copied back to a pinned
    host buffer asynchronously and handed out one step later (so the host never stalls the GPU).  Every step's
    inputs cross host->device and every step's result crosses device->host.
    """

    def __init__(self, trainer: "MVAETrainer", depth: int = 2):
        self.trainer = trainer
        self.dev = trainer.model.device_
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.d2h_stream = torch.cuda.Stream(device=self.dev)   # loss read-back: not between two steps on the main stream
        self.depth = depth
        self._slots = []

    def _slot(self, i, image, labels):
        while len(self._slots) <= i:
            self._slots.append(None)
        sl = self._slots[i]
        if sl is None or sl["x"].shape != image.shape or sl["x"].dtype != image.dtype:
            sl = {"x": torch.empty(image.shape, dtype=image.dtype, device=self.dev),
                  "y": torch.empty(labels.shape, dtype=labels.dtype, device=self.dev),
                  "ready": torch.cuda.Event(), "free": torch.cuda.Event(),
                  "loss_host": None, "loss_evt": torch.cuda.Event()}
            self._slots[i] = sl
        return sl

    def run(self, batches, **step_kwargs):
        """Generator over the per-step host loss arrays ([n_terms, 4] pinned float32 tensors)."""
        main = torch.cuda.current_stream(self.dev)
        it = iter(batches)
        pending = []           # (slot index) uploaded, waiting to be stepped
        inflight = []          # slots whose loss copy has been enqueued
        idx = 0

        def upload(k, batch):
            image, labels = batch
            sl = self._slot(k, image, labels)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(sl["free"])            # previous use of this slot has been consumed
                sl["x"].copy_(image, non_blocking=True)
                sl["y"].copy_(labels, non_blocking=True)
                sl["ready"].record(self.copy_stream)
            return sl

        nslots = self.depth + 1
        for _ in range(self.depth):
            b = next(it, None)
            if b is None:
                break
            pending.append(upload(idx % nslots, b))
            idx += 1
        while pending:
            sl = pending.pop(0)
            if getattr(self.trainer, "use_cuda_graph", False) or getattr(self.trainer, "dp_graph", False):
                # graph replay: the trainer stages the batch into its static buffers behind the upload's own event
                losses, _ = self.trainer.step(sl["x"], sl["y"], ready=sl["ready"], **step_kwargs)
            else:
                main.wait_event(sl["ready"])
                losses, _ = self.trainer.step(sl["x"], sl["y"], **step_kwargs)
            sl["free"].record(main)
            if sl["loss_host"] is None or sl["loss_host"].shape != losses.shape:
                sl["loss_host"] = torch.empty(losses.shape, dtype=losses.dtype).pin_memory()
            # (graph trainers hand out one of two alternating static loss buffers: this copy is complete - the generator
            # synchronises on it one step later - before the buffer's next step is enqueued)
            with torch.cuda.stream(self.d2h_stream):
                self.d2h_stream.wait_event(sl["free"])
                sl["loss_host"].copy_(losses, non_blocking=True)
                sl["loss_evt"].record(self.d2h_stream)
            losses.record_stream(self.d2h_stream)
            inflight.append(sl)
            b = next(it, None)
            if b is not None:
                pending.append(upload(idx % nslots, b))
                idx += 1
            if len(inflight) > 1:
                done = inflight.pop(0)
                done["loss_evt"].synchronize()
                yield done["loss_host"]
        for done in inflight:
            done["loss_evt"].synchronize()
            yield done["loss_host"]

