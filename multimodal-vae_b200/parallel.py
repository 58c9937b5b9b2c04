"""Data parallelism for the MVAE step: one process per GPU, torch.distributed (NCCL over NVLink) as plumbing.

The batch shards naturally: every rank runs the fused step on its own B/G samples with
per-replica BatchNorm statistics (DDP semantics), the flat fp32 gradient buffer - the whole model is ONE
contiguous buffer, 3.35 MB for MNIST - is summed with a single all-reduce, and the fused Adam kernel applies
grad_scale = 1/world_size.  Every loss term is a mean over the local shard, so averaging the gradients of equal
shards equals the gradient of the global mean.  There is no other exchange on the path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

from . import _lib
from .mnist import MVAE, MVAETrainer, _stream_ptr


def plan_buckets(sizes: Sequence[int], max_bucket: int) -> List[Tuple[int, int]]:
    """Greedy contiguous bucketing of a flat buffer made of tensors with `sizes` elements: [(start, stop)].
    A bucket closes when adding the next tensor would exceed max_bucket (a single larger tensor gets its own)."""
    out, start, cur = [], 0, 0
    for s in sizes:
        if cur > 0 and cur + s > max_bucket:
            out.append((start, start + cur))
            start, cur = start + cur, 0
        cur += s
    if cur > 0:
        out.append((start, start + cur))
    return out


def allreduce_flat_(flat: torch.Tensor, group=None, buckets: Sequence[Tuple[int, int]] = None) -> torch.Tensor:
    """Sum `flat` over the ranks in place (one collective per bucket; one bucket = the whole buffer by default)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return flat
    if not buckets:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    else:
        for a, b in buckets:
            dist.all_reduce(flat[a:b], op=dist.ReduceOp.SUM, group=group)
    return flat


def broadcast_model_(model: MVAE, src: int = 0, group=None) -> None:
    """Make every replica start from rank `src`'s parameters and buffers."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    dist.broadcast(model.flat_params, src, group=group)
    dist.broadcast(model.flat_buffers, src, group=group)
    dist.broadcast(model.flat_nbt, src, group=group)
    model.sync_low_precision()


class DataParallelTrainer(MVAETrainer):
    """MVAETrainer whose step is: local fused fwd+bwd -> all-reduce(flat grads) -> fused Adam(1/world)."""

    def __init__(self, model: MVAE, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 use_cuda_graph: bool = False, group=None, overlap: bool = True, fused: bool = True):
        super().__init__(model, lr=lr, betas=betas, eps=eps, use_cuda_graph=False)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.grad_scale = 1.0 / self.world      # step_masked's Adam: gradients are summed over the ranks first
        # every replica must draw its own reparametrisation noise: fold the rank into the Philox key
        model.noise_seed = (int(model.noise_seed) + 0x51ED27 * self.rank) & 0x7FFFFFFFFFFFFFFF
        self.dp_graph = use_cuda_graph
        self._dp_graphs = {}
        self.overlap = overlap
        self.dec_exchange_blocks = int(os.environ.get("MVAE_DP_DEC_BLOCKS", "32"))
        self.comm_stream = torch.cuda.Stream(device=model.device_)
        from .mnist import sizes
        self.enc_floats = int(sizes(model.n_latents, 2, model.dtype_code).encoder_param_floats)
        broadcast_model_(model, 0, group)
        # Gradient exchange: by default ONE fused kernel per bucket over NVLink peer memory (reduce-scatter + all-gather +
        # Adam, csrc/dp.cu); fused=False (or a failed symmetric-memory rendezvous) keeps the NCCL all-reduce + Adam kernel.
        self.fused = False
        if fused and self.world > 1 and self.world <= 8:
            try:
                self._setup_peer_memory()
                self.fused = True
            except Exception as exc:   # no peer access / symmetric memory unavailable: NCCL carries the gradients
                self.fused_error = repr(exc)

    def _setup_peer_memory(self) -> None:
        """Re-home the flat gradient buffer into symmetric memory and exchange the peer pointers."""
        import torch.distributed._symmetric_memory as symm_mem
        m = self.model
        grp = self.group if self.group is not None else dist.group.WORLD
        name = grp.group_name
        try:
            symm_mem.enable_symm_mem_for_group(name)   # older torch releases need the explicit opt-in
        except Exception:
            pass
        g = symm_mem.empty(m.flat_grads.numel(), dtype=torch.float32, device=m.device_)
        f = symm_mem.empty(64, dtype=torch.int32, device=m.device_)
        hg = symm_mem.rendezvous(g, name)
        hf = symm_mem.rendezvous(f, name)
        g.copy_(m.flat_grads)
        f.zero_()
        m.flat_grads = g
        for pname, kind, shape, off in m._table:
            if kind != 0:
                continue
            numel = 1
            for s_ in shape:
                numel *= s_
            block, _, idx, leaf = pname.split(".")
            getattr(getattr(getattr(m, block).net, idx), leaf).grad = g[off:off + numel].view(shape)
        self._peer_grads = [int(x) for x in hg.buffer_ptrs]
        self._peer_flags = [int(x) for x in hf.buffer_ptrs]
        self._symm = (g, f, hg, hf)
        torch.cuda.synchronize()
        dist.barrier(group=self.group)

    def _reduce_adam(self, lo: int, hi: int, blocks: int = 0) -> None:
        """Enqueue the fused exchange + update of bucket [lo, hi) on the current stream (a collective).  `blocks`: grid size
        (0 = the library's default)."""
        m, a = self.model, self.adam
        args = _lib.DpReduceAdamArgs()
        args.world, args.rank = self.world, self.rank
        for r in range(self.world):
            args.grads[r] = self._peer_grads[r]
            args.flags[r] = self._peer_flags[r]
        args.params = m.flat_params.data_ptr()
        args.adam_m, args.adam_v = a["m"].data_ptr(), a["v"].data_ptr()
        args.params_bf16 = None if m.flat_params_bf16 is None else m.flat_params_bf16.data_ptr()
        args.lo, args.hi = lo, hi
        args.lr, args.beta1, args.beta2, args.eps = a["lr"], a["betas"][0], a["betas"][1], a["eps"]
        args.grad_scale = 1.0 / self.world
        args.adam_step = m._adam_counter.data_ptr()
        args.blocks = int(blocks)
        _lib.check(_lib.load().mvae_dp_reduce_adam(C.byref(args), _stream_ptr()), "mvae_dp_reduce_adam")

    def _local_then_reduce(self, x, y, eps, terms, lambdas, annealing_factor, losses=None):
        """fwd + decoder backward -> [all-reduce(decoder bucket) on the comm stream || encoder backward]
        -> all-reduce(encoder bucket) -> Adam(1/world).  The flat parameter buffer is laid out as
        [encoders | decoders] exactly for this split (csrc/mnist_step.cu::make_layout)."""
        m = self.model
        tt, klw = self._norm(terms, x.shape[0], annealing_factor)
        split = self.enc_floats
        main = torch.cuda.current_stream(m.device_)
        total = m.flat_params.numel()
        if self.fused and self.world > 1:
            # forward + decoder-side backward; the decoder bucket is exchanged and updated on the communication stream
            # while the encoder-side backward runs; the encoder bucket follows on the main stream.  No NCCL, no separate Adam.
            out, _ = m._run(x, y, tt, lambdas, klw, eps=eps, backward=True, zero_grad=True, adam=None, losses=losses,
                            extra={"phase": 3, "advance_adam_step": 1})
            if self.overlap:
                self.comm_stream.wait_stream(main)
                with torch.cuda.stream(self.comm_stream):
                    # a small grid: the encoder-side backward beside it spreads over 65 SMs (csrc/chain.cu: column split, two
                    # parts per slab in phase 4) and a chain CTA needs an SM to itself - these blocks stay out of its way
                    # (2 GPUs, us/step: no split / 96 blocks 274; 2 parts / 19..48 blocks 267; 4 parts: 266..290)
                    self._reduce_adam(split, total, blocks=self.dec_exchange_blocks)
            m._run(x, y, tt, lambdas, klw, eps=eps, backward=True, zero_grad=False, adam=None, losses=losses,
                   extra={"phase": 4})
            if self.overlap:
                # the two calls share one set of flags: strictly one after the other on every rank
                main.wait_stream(self.comm_stream)
                self._reduce_adam(0, split)
            else:
                self._reduce_adam(0, total)
            return out
        if self.overlap and self.world > 1:
            out, _ = m._run(x, y, tt, lambdas, klw, eps=eps, backward=True, zero_grad=True, adam=None, losses=losses,
                            extra={"phase": 3, "advance_adam_step": 1})
            self.comm_stream.wait_stream(main)
            with torch.cuda.stream(self.comm_stream):
                allreduce_flat_(m.flat_grads[split:], self.group)
            m._run(x, y, tt, lambdas, klw, eps=eps, backward=True, zero_grad=False, adam=None, losses=losses,
                   extra={"phase": 4})
            allreduce_flat_(m.flat_grads[:split], self.group)
            main.wait_stream(self.comm_stream)
        else:
            out, _ = m._run(x, y, tt, lambdas, klw, eps=eps, backward=True, zero_grad=True, adam=None, losses=losses,
                            extra={"advance_adam_step": 1})
            allreduce_flat_(m.flat_grads, self.group)
        a = self.adam
        _lib.check(_lib.load().mvae_adam_step(
            C.c_void_p(m.flat_params.data_ptr()), C.c_void_p(m.flat_grads.data_ptr()), C.c_void_p(a["m"].data_ptr()),
            C.c_void_p(a["v"].data_ptr()),
            C.c_void_p(None if m.flat_params_bf16 is None else m.flat_params_bf16.data_ptr()),
            C.c_int64(m.flat_params.numel()), C.c_float(a["lr"]), C.c_float(a["betas"][0]), C.c_float(a["betas"][1]),
            C.c_float(a["eps"]), C.c_void_p(m._adam_counter.data_ptr()), C.c_float(1.0 / self.world), C.c_int(0),
            _stream_ptr()), "mvae_adam_step")
        return out

    def _reduce_gradients(self) -> None:
        """step_masked (inherited): the per-class backward passes accumulate locally, then the flat gradient buffer is
        summed over the ranks before the one Adam update (grad_scale = 1/world)."""
        allreduce_flat_(self.model.flat_grads, self.group)

    def step(self, image, text, eps=None, terms=("joint", "image", "text"), lambdas=((1.0, 1.0),) * 3,
             annealing_factor: float = 1.0, update: bool = True, outputs: bool = False, zero_grad: bool = True, ready=None):
        if not update or not zero_grad or outputs:
            raise NotImplementedError("DataParallelTrainer.step always zeroes the gradients, all-reduces and applies Adam; "
                                      "update=False / zero_grad=False / outputs=True are not supported (use MVAETrainer)")
        m = self.model
        y = text.to(m.device_, non_blocking=True).long().contiguous()
        if eps is not None:
            eps = eps.to(m.device_, torch.float32).contiguous()
        if not self.dp_graph:
            return self._local_then_reduce(m.to_act(image), y, eps, terms, lambdas, annealing_factor), None
        x = self._graph_input(image)
        a_ = self.adam
        self._slot ^= 1
        key = (x.shape[0], tuple(terms), tuple(map(tuple, lambdas)), float(annealing_factor), eps is not None,
               float(a_["lr"]), tuple(map(float, a_["betas"])), float(a_["eps"]), x.dtype, self._slot)
        ent = self._graph_cache_get(self._dp_graphs, key)
        if ent is None:
            # uint8 pixels are converted by the staging copy (MVAETrainer._stage), not by a node of the graph
            sx = torch.empty(x.shape, device=m.device_, dtype=m.act_dtype()) if x.dtype == torch.uint8 else torch.empty_like(x)
            sy = torch.empty_like(y)
            se = torch.empty_like(eps) if eps is not None else None
            losses = torch.empty(len(terms), 4, device=m.device_, dtype=torch.float32)
            # NCCL must have been used once outside capture (communicator setup is not capturable)
            allreduce_flat_(torch.zeros(8, device=m.device_), self.group)
            m.workspace(x.shape[0])
            torch.cuda.synchronize()
            lib = _lib.load()
            before = lib.mvae_launch_count()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._local_then_reduce(sx, sy, se, terms, lambdas, annealing_factor, losses=losses)
            ent = {"graph": graph, "x": sx, "y": sy, "eps": se, "losses": losses, "fresh": True,
                   "free": torch.cuda.Event(), "ready": torch.cuda.Event(),
                   "launches": int(lib.mvae_launch_count() - before) + (1 if x.dtype == torch.uint8 else 0)}
            self._graph_cache_put(self._dp_graphs, key, ent)
        self._stage(ent, x, y, eps, ready)
        ent["graph"].replay()
        ent["free"].record(torch.cuda.current_stream(m.device_))
        self.last_graph_launches = ent["launches"] + 1
        return ent["losses"], None
