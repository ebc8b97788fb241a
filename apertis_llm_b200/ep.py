"""Expert parallelism for AdaptiveExpertSystem: experts sharded across the ranks of one node, token rows
exchanged with an all-to-all over NCCL (NVLink 5 / NVSwitch), batch data-parallel everywhere else.

The reference has no expert parallelism (only DDP, pipeline.py:435-466); this is the B200-native addition
SURVEY.md section 8(e) specifies.  Semantics are those of replicated experts under DDP:

* every rank routes ITS OWN tokens and applies the capacity limit to its local S (core.py:508-511 uses the
  local S), so the forward is identical to running all experts locally;
* rank r owns experts [r*E/W, (r+1)*E/W).  The local permuted layout uses FIXED expert segments of
  `seg = round_up(cap, 128)` rows (no counts need to reach the host: the exchange has equal splits), so
  `xn[W, El*seg, Dm]` goes out with one all_to_all_single, the owner runs the grouped GEMMs over the
  received `[W, El*seg, Dm]` rows (row tile -> local expert is a static map), and the result comes back the
  same way.  The per-expert LayerNorm affine is all-gathered (E*Dm floats) and applied at the source, so the
  numerics equal the single-GPU path bit for bit;
* backward mirrors it (dY out, dXn back); expert gradients are sums over all ranks' tokens and are scaled by
  1/W so that they equal what DDP's gradient averaging would give for replicated experts.

Two transports move the rows (APERTIS_B200_EP = peer | nccl | auto, default auto):

* ``peer`` (bf16 path, all ranks on one NVLink / NVSwitch node, torch symmetric memory available): no all-to-all at all.
  Every rank owns four peer-mapped buffers ``[W, El*seg, Dm]`` and every exchange is fused into the kernel that PRODUCES
  the rows, which stores them straight into the consuming rank's buffer: permute + LayerNorm into the owners' receive
  buffers (``ab_ep_permute_ln``), the second expert GEMM's epilogue into the source ranks' buffers
  (``ab_ep_grouped_gemm_nt``), and in the backward the un-permutation's dY rows (``ab_ep_unpermute_bwd``) and the input-
  gradient GEMM's epilogue (``ab_ep_grouped_gemm_nn``).  The NVLink transfer overlaps the producing kernel tile by tile;
  what remains between the ranks are five ~7 us barriers per step (``_SymmetricMemory.barrier``): all producers of a
  buffer before its consumers.
* ``nccl``: ``all_to_all_single`` between the same kernels (any backend NCCL supports; also the fp32-parity mode).
"""
from __future__ import annotations

import os
from typing import List

import ctypes

import torch
import torch.distributed as dist

from . import _lib, ops
from ._lib import ROW_ALIGN, call, dt, ptr, query, stream_ptr


# ------------------------------------------------------------------------------------------------
# layout helpers (pure, device agnostic: covered by the CPU gloo tests)
# ------------------------------------------------------------------------------------------------
def segment_rows(cap: int) -> int:
    """Rows reserved per expert in the exchange layout."""
    return (int(cap) + ROW_ALIGN - 1) // ROW_ALIGN * ROW_ALIGN


def recv_tile_expert(W: int, El: int, seg: int, device=None) -> torch.Tensor:
    """Local expert of every 128-row tile of the received buffer [W, El*seg, C] (source-major)."""
    tiles_per_seg = seg // ROW_ALIGN
    return torch.arange(El, dtype=torch.int32, device=device).repeat_interleave(tiles_per_seg).repeat(W)


def local_seg_off(El: int, seg: int, device=None) -> torch.Tensor:
    return torch.arange(El + 1, dtype=torch.int32, device=device) * seg


def all_to_all_equal(send: torch.Tensor, group) -> torch.Tensor:
    """send [W, n, C] (block d goes to rank d) -> recv [W, n, C] (block s came from rank s)."""
    W = dist.get_world_size(group)
    assert send.shape[0] == W and send.is_contiguous()
    recv = torch.empty_like(send)
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(recv.view(-1), send.view(-1), group=group)
        return recv
    # gloo has no all_to_all: pairwise exchange (used by the CPU tests of the host logic)
    rank = dist.get_rank(group)
    recv[rank].copy_(send[rank])
    ops_ = []
    for peer in range(W):
        if peer == rank:
            continue
        ops_.append(dist.P2POp(dist.isend, send[peer], dist.get_global_rank(group, peer), group))
        ops_.append(dist.P2POp(dist.irecv, recv[peer], dist.get_global_rank(group, peer), group))
    for w in dist.batch_isend_irecv(ops_):
        w.wait()
    return recv


def all_to_all_equal_start(send: torch.Tensor, group):
    """Same exchange, started asynchronously on NCCL's stream: returns (recv, wait) and the caller launches independent
    kernels before calling wait() (other backends: exchanged synchronously, wait is a no-op)."""
    if dist.get_backend(group) != "nccl":
        return all_to_all_equal(send, group), (lambda: None)
    assert send.shape[0] == dist.get_world_size(group) and send.is_contiguous()
    recv = torch.empty_like(send)
    work = dist.all_to_all_single(recv.view(-1), send.view(-1), group=group, async_op=True)
    return recv, work.wait


def all_gather_cat(t: torch.Tensor, group) -> torch.Tensor:
    W = dist.get_world_size(group)
    out = torch.empty((W,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1), group=group) if dist.get_backend(group) == "nccl" else \
        dist.all_gather(list(out.unbind(0)), t.contiguous(), group=group)
    return out.view((-1,) + tuple(t.shape[1:]))


# ------------------------------------------------------------------------------------------------
# peer-memory transport
# ------------------------------------------------------------------------------------------------
_EP_MODE = os.environ.get("APERTIS_B200_EP", "auto")

_peer_cache = {}          # (group name, rank, rows, Dm) -> _PeerState
_peer_failed = False
_warned_reuse = False


class _PeerState:
    """The four peer-mapped row buffers of one (group, shape) and their address tables."""

    def __init__(self, group, rows: int, Dm: int, device):
        import torch.distributed._symmetric_memory as symm_mem
        self.W = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.rows, self.Dm = rows, Dm
        names = ("xn", "y", "dy", "dxn")
        self.buf, self.hdl, self.ptrs = {}, {}, {}
        for n in names:
            t = symm_mem.empty(rows, Dm, dtype=torch.bfloat16, device=device)
            h = symm_mem.rendezvous(t, group)
            self.buf[n], self.hdl[n] = t, h
            self.ptrs[n] = (ctypes.c_uint64 * self.W)(*[int(p) for p in h.buffer_ptrs])
        self.sync = self.hdl["xn"]
        self.version = 0          # forwards run on these buffers (a backward must see the version its forward left)
        self.side = torch.cuda.Stream(device=device)      # the backward's last barrier waits beside the weight-gradient GEMMs

    def barrier(self, channel: int = 0):
        """All ranks' preceding work on the current stream is complete and visible (a kernel: CUDA-graph capturable)."""
        self.sync.barrier(channel=channel)


_uniform_checked = set()
_recv_plans = {}


def _recv_plan(W: int, El: int, seg: int, device):
    """The owner-side plan of the received layout (static for a given world size, local expert count and segment size):
    built once, not by a handful of tiny kernels every step."""
    key = (W, El, seg, device.index)
    p = _recv_plans.get(key)
    if p is None:
        p = dict(tile_expert=recv_tile_expert(W, El, seg, device),
                 n_rows=torch.full((2,), W * El * seg, dtype=torch.int32, device=device),
                 seg_off=local_seg_off(El, seg, device))
        if not torch.cuda.is_current_stream_capturing():
            _recv_plans[key] = p
    return p


def _check_uniform(group, S: int, cap: int, training: bool, device):
    """The exchange has equal splits: every rank must bring the same number of tokens and the same capacity.  Checked with
    one small all-gather the first time a (tokens, capacity, mode) combination is seen; ragged batches fail here, loudly,
    instead of hanging NCCL or exchanging misaligned rows."""
    key = (id(group), S, cap, training)
    if key in _uniform_checked:
        return
    mine = torch.tensor([S, cap, int(training)], dtype=torch.int64, device=device)
    allv = [torch.empty_like(mine) for _ in range(dist.get_world_size(group))]
    dist.all_gather(allv, mine, group=group)
    vals = torch.stack(allv).cpu()
    if not bool((vals == vals[0]).all()):
        raise RuntimeError("apertis_b200 EP: every rank of the expert-parallel group must process the same number of tokens with the "
                           f"same capacity and mode; got (tokens, capacity, training) per rank = {vals.tolist()}")
    _uniform_checked.add(key)


def _peer_state(group, rows: int, Dm: int, device, owner=0):
    """Buffers of one layer (`owner`) and shape: a layer's receive buffer doubles as its saved activation, so layers do
    not share them."""
    global _peer_failed
    if _EP_MODE == "nccl" or _peer_failed or dist.get_backend(group) != "nccl":
        return None
    key = (id(group), dist.get_rank(group), rows, Dm, device.index, owner)
    st = _peer_cache.get(key)
    if st is None:
        if torch.cuda.is_current_stream_capturing():
            return None                       # buffers are created eagerly (warm-up) and reused by captured steps
        try:
            st = _PeerState(group, rows, Dm, device)
        except Exception as ex:               # no symmetric memory on this platform: NCCL transport
            if _EP_MODE == "peer":
                raise
            _peer_failed = True
            import warnings
            warnings.warn(f"apertis_b200 EP: peer-memory transport unavailable ({type(ex).__name__}: {ex}); using NCCL all-to-all")
            return None
        _peer_cache[key] = st
    return st


# ------------------------------------------------------------------------------------------------
# the autograd node
# ------------------------------------------------------------------------------------------------
class _MoEExpertsEP(torch.autograd.Function):
    """ops._MoEExperts with the expert MLPs executed on the owning ranks."""

    @staticmethod
    @ops._on_tensor_device
    def forward(ctx, x2, rn_w, rn_b, Wr, br, noise, noise_scale, ln_w, ln_b, W1, b1, W2, b2, res, cfg, group):
        _lib.ensure_device(x2.device)
        dev = x2.device
        W = dist.get_world_size(group)
        S, Dm = x2.shape
        El, I, _ = W1.shape
        E = El * W
        K, act, training, precise = cfg["K"], cfg["act"], cfg["training"], cfg["precise"]
        x2 = x2.contiguous()
        f = lambda t: t.float().contiguous()
        rn_w, rn_b, Wr, br, b1, b2, W1, W2 = map(f, (rn_w, rn_b, Wr, br, b1, b2, W1, W2))
        # every source rank normalises the rows it dispatches: all experts' LayerNorm parameters, one collective for both,
        # repeated only when the parameters have changed (in-place optimizer updates bump the version counters; every rank
        # sees the same sequence of updates, so all ranks gather, or reuse, together)
        ln_key = (ln_w._version, ln_b._version, ln_w.data_ptr(), ln_b.data_ptr(), E, Dm)
        ln_cache = cfg.get("_ln_cache")
        if ln_cache is not None and ln_cache.get("key") == ln_key:
            ln_w_full, ln_b_full = ln_cache["val"]
        else:
            ln_wb = all_gather_cat(torch.stack([f(ln_w), f(ln_b)]).unsqueeze(0), group)           # [W, 2, El, Dm]
            ln_w_full = ln_wb[:, 0].reshape(E, Dm).contiguous()   # [E, Dm]
            ln_b_full = ln_wb[:, 1].reshape(E, Dm).contiguous()
            if ln_cache is not None:
                ln_cache["key"], ln_cache["val"] = ln_key, (ln_w_full, ln_b_full)
        use_noise = noise is not None and noise_scale is not None
        r = ops.moe_route(x2, rn_w, rn_b, cfg["eps"], Wr, br, f(noise) if use_noise else None,
                          f(noise_scale) if use_noise else None, K, cfg.get("quant", _lib.ROUTER_EXACT))
        _check_uniform(group, S, cfg["cap"], training, dev)
        if cfg["cap"] >= S:
            # no capacity limit (evaluation, or use_expert_capacity_limit=False): a segment only has to hold the fullest
            # expert of any rank, not all S tokens - one MAX all-reduce and a host read, off the training path
            cnt = torch.bincount(r["idx"].reshape(-1).long(), minlength=E)[:E]
            if cfg["active"] is not None:
                cnt = cnt * cfg["active"].to(cnt.dtype)
            mx = cnt.max().reshape(1)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
            seg = segment_rows(max(1, int(mx.item())))
        else:
            seg = segment_rows(min(cfg["cap"], S))
        plan = ops.moe_plan(r["idx"], r["w"], E, cfg["cap"], cfg["active"], fixed_seg=seg)
        rows_local = E * seg                                  # == W * El * seg
        cdt = torch.float32 if precise else torch.bfloat16
        peer = None if precise else _peer_state(group, rows_local, Dm, dev, cfg.get("_owner", 0))
        rank = dist.get_rank(group)
        rplan = _recv_plan(W, El, seg, dev)
        if peer is not None:
            # ---- dispatch fused into permute + LayerNorm: every row goes straight into its owner's receive buffer
            peer.version += 1
            peer.barrier(0)                   # the owners are done with last step's rows
            call("ab_ep_permute_ln", ptr(x2), ptr(r["stats"]), ptr(ln_w_full), ptr(ln_b_full), ptr(plan["tok_of_row"]),
                 ptr(plan["tile_expert"]), ptr(plan["n_rows"]), peer.ptrs["xn"], W, rank, El * seg, Dm, ROW_ALIGN, rows_local,
                 dt(x2), dt(cdt), stream_ptr())
            w1, w2 = ops._cast_bf16(W1), ops._cast_bf16(W2)      # before the barrier: absorbs the ranks' skew
            peer.barrier(1)                   # every source has written its rows
            xr = peer.buf["xn"]
        else:
            xn = torch.empty(rows_local, Dm, dtype=cdt, device=dev)
            call("ab_moe_permute_ln", ptr(x2), ptr(r["stats"]), ptr(ln_w_full), ptr(ln_b_full), ptr(plan["tok_of_row"]),
                 ptr(plan["tile_expert"]), ptr(plan["n_rows"]), ptr(xn), Dm, ROW_ALIGN, rows_local, dt(x2), dt(cdt), stream_ptr())
            # ---- dispatch
            xr_w, xr_wait = all_to_all_equal_start(xn.view(W, El * seg, Dm), group)
            # while the rows travel: the bf16 weight shadows
            w1 = ops._split_cols(W1.view(El * I, Dm), 1) if precise else ops._cast_bf16(W1)
            w2 = ops._split_cols(W2.view(El * Dm, I), 1) if precise else ops._cast_bf16(W2)
            xr_wait()
            xr = xr_w.view(rows_local, Dm)
        if precise:
            a1, k1 = ops._split_cols(xr, 0), 3 * Dm
        else:
            a1, k1 = xr, Dm
        drop_p = float(cfg.get("drop_p", 0.0)) if training else 0.0
        drop_seed = torch.randint(0, 2 ** 31 - 1, (2,), device=dev, dtype=torch.int32) if drop_p > 0.0 else None
        h, hpre = ops.grouped_gemm("nt", a1, w1, rplan, I, k1, El, bias=b1, epi=_lib.EPI_BIAS_ACT, act=act, out_dtype=cdt, want_c2=True,
                                   drop_p=drop_p, drop_seed=drop_seed)
        if precise:
            a2, k2 = ops._split_cols(h, 0), 3 * I
        else:
            a2, k2 = h, I
        # the caller's output dropout and residual add (core.py:918-919) happen inside the combine kernel when handed in
        out_p = float(cfg.get("out_drop_p", 0.0)) if training else 0.0
        out_seed = torch.randint(0, 2 ** 31 - 1, (2,), device=dev, dtype=torch.int32) if out_p > 0.0 else None
        resc = res.reshape(S, Dm).float().contiguous() if res is not None else None
        out = torch.empty(S, Dm, dtype=x2.dtype if res is None else torch.float32, device=dev)
        if peer is not None:
            # ---- combine fused into the second GEMM: its epilogue stores every row into the source rank's buffer
            call("ab_ep_grouped_gemm_nt", ptr(a2), ptr(w2), ptr(b2), None, peer.ptrs["y"], W, rank, El * seg, ptr(rplan["tile_expert"]),
                 ptr(rplan["n_rows"]), rows_local, Dm, k2, El, _lib.EPI_BIAS, act, dt(cdt), stream_ptr())
            peer.barrier(2)                   # every owner has delivered its expert outputs
            y = peer.buf["y"]                 # this rank's rows, in its own permuted layout; kept for the backward
            call("ab_moe_unpermute", ptr(y), ptr(plan["row_of"]), ptr(r["w"]), ptr(resc), ptr(out), out_p, ptr(out_seed), S, K, Dm, dt(y), dt(out),
                 stream_ptr())
        else:
            yr = ops.grouped_gemm("nt", a2, w2, rplan, Dm, k2, El, bias=b2, epi=_lib.EPI_BIAS, out_dtype=cdt)
            # ---- combine
            y = all_to_all_equal(yr.view(W, El * seg, Dm), group).view(rows_local, Dm)
            call("ab_moe_unpermute", ptr(y), ptr(plan["row_of"]), ptr(r["w"]), ptr(resc), ptr(out), out_p, ptr(out_seed), S, K, Dm, dt(y), dt(out),
                 stream_ptr())
        aux = r["aux"]
        zero = torch.zeros((), dtype=x2.dtype, device=dev)
        lb = (cfg["lb_coef"] * E / (S * S)) * (aux[:E] * aux[E:2 * E]).sum() if (training and cfg["lb_coef"] > 0) else zero
        rz = (cfg["rz_coef"] / S) * aux[2 * E] if (training and cfg["rz_coef"] > 0) else zero
        ctx.cfg = dict(cfg, S=S, Dm=Dm, E=E, El=El, I=I, W=W, seg=seg, use_noise=use_noise, rows=rows_local, cdt=cdt, drop_p=drop_p,
                       out_p=out_p, has_res=res is not None, res_shape=res.shape if res is not None else None,
                       res_dtype=res.dtype if res is not None else None)
        ctx.drop_seed = drop_seed
        ctx.out_seed = out_seed
        ctx.shadows = None if precise else (w1, w2)
        ctx.group = group
        ctx.peer = peer
        ctx.peer_version = peer.version if peer is not None else 0
        ctx.plan = {k: v for k, v in plan.items() if torch.is_tensor(v)}
        ctx.rplan = rplan
        ctx.save_for_backward(x2, rn_w, rn_b, Wr, br, ln_w_full, W1, W2, noise if use_noise else None, r["stats"], r["gates"],
                              r["idx"], r["probs"], r["lse"], r["lclean"], r["w"], aux, xr, h, hpre, y)
        counts = plan["counts"]
        ctx.mark_non_differentiable(counts)
        cfg["_routing"] = (r["idx"], plan["row_of"])
        return out, lb.to(x2.dtype), rz.to(x2.dtype), counts

    @staticmethod
    @ops._on_tensor_device
    def backward(ctx, dout, dlb, drz, _dcounts):
        (x2, rn_w, rn_b, Wr, br, ln_w_full, W1, W2, noise, stats, gates, idx, probs, lse, lclean, w, aux, xr, h, hpre, y) = ctx.saved_tensors
        cfg, plan, rplan, group = ctx.cfg, ctx.plan, ctx.rplan, ctx.group
        S, Dm, E, El, I, K, W, seg = (cfg[k] for k in ("S", "Dm", "E", "El", "I", "K", "W", "seg"))
        act, precise, rows, cdt = cfg["act"], cfg["precise"], cfg["rows"], cfg["cdt"]
        dev = x2.device
        f32 = dict(dtype=torch.float32, device=dev)
        dout = dout.contiguous()
        peer = ctx.peer
        rank = dist.get_rank(group)
        dw_row = torch.empty(rows, **f32)
        if peer is not None and ctx.peer_version != peer.version:
            # Another forward of this layer has rewritten the peer-mapped buffers that double as saved activations.  Under
            # activation checkpointing (the reference trainer's default, core.py:1258-1272) that forward is the
            # recomputation and wrote the same rows again: fine.  Any other forward-forward-backward-backward schedule on
            # one layer would read the later forward's rows here.
            msg = ("apertis_b200 EP (peer transport): a later forward of this layer ran before this backward.  Expected under "
                   "activation checkpointing (the recomputation rewrites identical rows); any other schedule that keeps two "
                   "forwards of one layer in flight needs APERTIS_B200_EP=nccl.")
            if os.environ.get("APERTIS_B200_EP_STRICT", "0") == "1":
                raise RuntimeError(msg)
            global _warned_reuse
            if not _warned_reuse:
                _warned_reuse = True
                import warnings
                warnings.warn(msg)
        if peer is not None:
            # dY rows go straight into their owners' receive buffers
            call("ab_ep_unpermute_bwd", ptr(dout), ptr(y), ptr(w), ptr(plan["tok_of_row"]), ptr(plan["slot_of_row"]), ptr(plan["n_rows"]),
                 peer.ptrs["dy"], W, rank, El * seg, ptr(dw_row), cfg["out_p"], ptr(ctx.out_seed), K, Dm, rows, dt(dout), dt(y), dt(cdt),
                 stream_ptr())
            peer.barrier(3)
            dyr = peer.buf["dy"]
        else:
            dy = torch.empty(rows, Dm, dtype=cdt, device=dev)
            call("ab_moe_unpermute_bwd", ptr(dout), ptr(y), ptr(w), ptr(plan["tok_of_row"]), ptr(plan["slot_of_row"]), ptr(plan["n_rows"]),
                 ptr(dy), ptr(dw_row), cfg["out_p"], ptr(ctx.out_seed), K, Dm, rows, dt(dout), dt(y), dt(cdt), stream_ptr())
            dyr = all_to_all_equal(dy.view(W, El * seg, Dm), group).view(rows, Dm)
        lseg = rplan["seg_off"]
        stride = El * seg
        # bias gradients (column sums over each local expert's rows) on the auxiliary stream, launched after the GEMMs
        cmain, cside = torch.cuda.current_stream(dev), ops._side_stream(dev)
        db2 = torch.empty(El, Dm, **f32)
        db1 = torch.empty(El, I, **f32)
        ws_b2 = ops._u8(query("ab_moe_segment_colsum_workspace_bytes", Dm, ROW_ALIGN, rows), dev)
        ws_b1 = ops._u8(query("ab_moe_segment_colsum_workspace_bytes", I, ROW_ALIGN, rows), dev)
        if precise:
            # row-stacked splits: every (source, local expert) block of `seg` rows is one group
            G = W * El
            lseg3 = (lseg * 3).contiguous()
            w2r = ops._split_rows(W2.view(El * Dm, I), 1, None, El, Dm)
            dhpre = ops.grouped_gemm("nn", ops._split_cols(dyr, 0), w2r, rplan, I, 3 * Dm, El, aux=hpre, epi=_lib.EPI_DACT, act=act, out_dtype=cdt,
                                     drop_p=cfg["drop_p"], drop_seed=ctx.drop_seed)
            sr = lambda t, which: ops._split_rows(t, which, None, G, seg)
            w1r = ops._split_rows(W1.view(El * I, Dm), 1, None, El, I)
            dxnr = ops.grouped_gemm("nn", ops._split_cols(dhpre, 0), w1r, rplan, Dm, 3 * I, El, out_dtype=torch.float32)
            # the gradient rows travel back to their source ranks while the weight gradients are computed
            dxn_w, dxn_wait = all_to_all_equal_start(dxnr.view(W, El * seg, Dm), group)
            dW2 = ops.grouped_gemm_tn(sr(dyr, 0), sr(h, 1), lseg3, Dm, I, El, nsrc=W, src_stride=3 * stride)
            dW1 = ops.grouped_gemm_tn(sr(dhpre, 0), sr(xr, 1), lseg3, I, Dm, El, nsrc=W, src_stride=3 * stride)
        else:
            w1b, w2b = ctx.shadows
            dhpre = ops.grouped_gemm("nn", dyr, w2b, rplan, I, Dm, El, aux=hpre, epi=_lib.EPI_DACT, act=act, out_dtype=cdt,
                                     drop_p=cfg["drop_p"], drop_seed=ctx.drop_seed)
            if peer is not None:
                # the input-gradient GEMM's epilogue stores every row into its source rank's buffer; the barrier that
                # completes the exchange waits on a side stream while the weight gradients are computed
                call("ab_ep_grouped_gemm_nn", ptr(dhpre), ptr(w1b), None, None, peer.ptrs["dxn"], W, rank, El * seg, ptr(rplan["tile_expert"]),
                     ptr(rplan["n_rows"]), rows, Dm, I, El, _lib.EPI_NONE, act, dt(torch.bfloat16), stream_ptr())
                main = torch.cuda.current_stream(dev)
                peer.side.wait_stream(main)
                with torch.cuda.stream(peer.side):
                    peer.barrier(4)
                dxn_w = peer.buf["dxn"]
                dxn_wait = lambda: main.wait_stream(peer.side)
            else:
                dxnr = ops.grouped_gemm("nn", dhpre, w1b, rplan, Dm, I, El, out_dtype=torch.bfloat16)
                dxn_w, dxn_wait = all_to_all_equal_start(dxnr.view(W, El * seg, Dm), group)
            dW2 = ops.grouped_gemm_tn(dyr, h, lseg, Dm, I, El, nsrc=W, src_stride=stride)
            dW1 = ops.grouped_gemm_tn(dhpre, xr, lseg, I, Dm, El, nsrc=W, src_stride=stride)
        # after the GEMMs (which fill every SM): the column sums run beside the latency-bound tail of the node
        ops._colsum_side(cside, cmain, dyr, rplan, db2, ws_b2, Dm, El, rows, dev)
        ops._colsum_side(cside, cmain, dhpre, rplan, db1, ws_b1, I, El, rows, dev)
        ws = ops._u8(query("ab_moe_permute_ln_bwd_workspace_bytes", Dm, ROW_ALIGN, rows), dev)
        # ---- gradient rows are back on their source ranks
        dxn_wait()
        dxn = dxn_w.view(rows, Dm)
        dxrow = torch.empty(rows, Dm, **f32)
        dln_w = torch.empty(E, Dm, **f32)
        dln_b = torch.empty(E, Dm, **f32)
        call("ab_moe_permute_ln_bwd", ptr(dxn), ptr(x2), ptr(stats), ptr(ln_w_full), ptr(plan["tok_of_row"]), ptr(plan["tile_expert"]),
             ptr(plan["n_rows"]), ptr(dxrow), ptr(dln_w), ptr(dln_b), ptr(ws), ws.numel(), Dm, E, ROW_ALIGN, rows, dt(x2), dt(dxn),
             stream_ptr())
        dln = torch.stack([dln_w, dln_b])                      # [2, E, Dm]: sum over source ranks, keep the local experts
        dln_work = dist.all_reduce(dln, group=group, async_op=True)        # overlaps the router backward below
        inv = 1.0 / W
        training = cfg["training"]
        g_lb = (dlb.float() * (cfg["lb_coef"] * E / S)) if (training and cfg["lb_coef"] > 0) else torch.zeros((), **f32)
        g_rz = (drz.float() * (cfg["rz_coef"] / S)) if (training and cfg["rz_coef"] > 0) else torch.zeros((), **f32)
        scal = torch.stack([g_lb.reshape(()), g_rz.reshape(())]).contiguous()
        fvec = (aux[E:2 * E] / S).contiguous()
        dx = torch.empty_like(x2)
        dWr, dbr = torch.empty(E, Dm, **f32), torch.empty(E, **f32)
        drn_w, drn_b = torch.empty(Dm, **f32), torch.empty(Dm, **f32)
        dns = torch.empty(E, **f32) if cfg["use_noise"] else None
        nws = query("ab_moe_router_bwd_workspace_bytes", S, Dm, E)
        ws2 = ops._u8(nws, dev)
        call("ab_moe_router_bwd", ptr(x2), ptr(stats), ptr(rn_w), ptr(rn_b), ptr(Wr), ptr(br), ptr(gates), ptr(idx), ptr(probs),
             ptr(lse), ptr(lclean), ptr(noise.float().contiguous()) if cfg["use_noise"] else None, ptr(fvec), ptr(scal), ptr(dw_row),
             ptr(dxrow), ptr(plan["row_of"]), ptr(dx), ptr(dWr), ptr(dbr), ptr(drn_w), ptr(drn_b), ptr(dns), ptr(ws2), ws2.numel(),
             S, Dm, E, K, dt(x2), stream_ptr())
        dln_work.wait()
        dln_w_l = dln[0, rank * El:(rank + 1) * El] * inv
        dln_b_l = dln[1, rank * El:(rank + 1) * El] * inv
        dres = dout.reshape(cfg["res_shape"]).to(cfg["res_dtype"]) if cfg["has_res"] else None       # the residual passes the gradient on
        cmain.wait_stream(cside)                                     # bias gradients
        return (dx, drn_w, drn_b, dWr, dbr, None, dns, dln_w_l, dln_b_l, dW1 * inv, db1 * inv, dW2 * inv, db2 * inv, dres, None, None)


def moe_experts_ep(module, x2, noise, noise_scale, cfg, res=None):
    """Entry used by AdaptiveExpertSystem.forward when an expert-parallel group is set.  res (optional, [S, Dm]): the
    caller's residual; with it the result is res + dropout(cfg['out_drop_p'])(moe(x2))."""
    cfg["_owner"] = id(module)
    cfg["_ln_cache"] = module.__dict__.setdefault("_ep_ln_cache", {})
    return _MoEExpertsEP.apply(x2, module.router_norm.weight, module.router_norm.bias, module.router.weight, module.router.bias,
                               noise, noise_scale, module.expert_ln_weight, module.expert_ln_bias, module.expert_w1,
                               module.expert_b1, module.expert_w2, module.expert_b2, res, cfg, module.ep_group)


def replicated_parameters(layer: torch.nn.Module) -> List[torch.nn.Parameter]:
    """Parameters that are replicated across the EP group (everything but the sharded expert tensors);
    these are the ones a DDP wrapper (pipeline.py:463) must all-reduce."""
    return [p for n, p in layer.named_parameters() if ".expert_" not in n and not n.startswith("expert_")]
