"""Multi-GPU partitioning of the hot path (one process per GPU, torch.distributed; NCCL on GPUs, gloo in the
CPU tests).  The reference has no distributed code at all; both schemes below are exact.

* batch sharding  -- reads are independent through every op on the path.  Inference needs no collective; training
  needs one gradient all-reduce (SUM) per step, with the batch-MEAN cross-entropy term pre-scaled by 1/world and
  the batch-SUM CTC term left alone (legacy_code/train.py:39 vs :46).
* time sharding   -- the networks are fixed-receptive-field stacks: a rank that owns frames [s, e) of a long read
  needs `halo_left` / `halo_right` extra input samples from its neighbours, recomputes that fringe and keeps only
  its own span.  Zero padding at the TRUE ends of the read is what the kernels do natively (TMA / bounds-check
  zero fill), so the first and last rank simply get no halo on that side.
"""
import torch

from .functional import tap_offsets


# ----------------------------------------------------------------------------------------- batch sharding
def shard_range(n, rank, world):
    """Contiguous split of n items: -> (start, end) of `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def joint_loss_for_backward(xe_local_sum_over_t, ctc_local_sum, T, T_ctc, world):
    """Local loss whose gradient, all-reduced with SUM, equals the reference's single-process gradient of
    xe/T + ctc/T' (legacy_code/train.py:50-54).  `xe_local_sum_over_t` is already a mean over the LOCAL batch."""
    return xe_local_sum_over_t / (T * world) + ctc_local_sum / T_ctc


def allreduce_gradients(params, group=None, bucket_bytes=256 << 20):
    """Bucketed SUM all-reduce of .grad over the process group (NCCL over NVLink on GPUs).  Buckets are filled in
    reverse parameter order -- the order backward produces them.  Each bucket is ONE flat tensor: the gradients are
    gathered into it with a fused multi-tensor copy, reduced in place, and the parameters' .grad are re-pointed at
    views of it (no copy back) -- 442 parameters cost 3 launches instead of ~900."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    plist = [p for p in reversed(list(params)) if p.grad is not None]
    n_buckets, i = 0, 0
    while i < len(plist):
        bucket, size = [], 0
        dt, dev = plist[i].grad.dtype, plist[i].grad.device
        while i < len(plist) and plist[i].grad.dtype == dt and (not bucket or size < bucket_bytes):
            bucket.append(plist[i])
            size += plist[i].grad.numel() * plist[i].grad.element_size()
            i += 1
        flat = torch.empty(sum(p.grad.numel() for p in bucket), dtype=dt, device=dev)
        views = list(flat.split([p.grad.numel() for p in bucket]))
        srcs = [p.grad.reshape(-1) for p in bucket]
        if hasattr(torch, "_foreach_copy_"):
            torch._foreach_copy_(views, srcs)
        else:  # pragma: no cover
            for v, g in zip(views, srcs):
                v.copy_(g)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        for p, v in zip(bucket, views):
            p.grad = v.view_as(p)
        n_buckets += 1
    return n_buckets


# ----------------------------------------------------------------------------------------- receptive fields
def stack_halo(tap_lists):
    """tap_lists: one list of frame offsets per layer -> (left, right) input frames a frame depends on."""
    left = sum(max(0, -min(offs)) for offs in tap_lists)
    right = sum(max(0, max(offs)) for offs in tap_lists)
    return left, right


def wavenet_halo(model):
    """Config-2 WaveNet (entry k=2 + 2 x (1..512)): (2047, 0)."""
    taps = [tap_offsets(model.entry_kwidth, 1, True)]
    taps += [tap_offsets(k, d, True) for (_ci, _co, k, d) in model.layers]
    return stack_halo(taps)


def raw_ctcnet_halo(model):
    """RawCTCNet: featuriser (reads x[t'-(fk-1) .. t']) + input block + stack.  ecoli config, fk=3: (51, 45)."""
    fk = model.feature_kwidth
    taps = [[j - (fk - 1) for j in range(fk)]]
    taps.append(tap_offsets(model.input_kernel_size, model.input_dilation, model.causal))
    taps += [tap_offsets(k, d, model.causal) for (_ci, _co, k, d) in model.layers]
    return stack_halo(taps)


# ----------------------------------------------------------------------------------------- time sharding
def time_shard_plan(T, rank, world, halo_left, halo_right, align=1):
    """-> dict(start, end, lo, hi): this rank owns input frames [start, end) and must read [lo, hi)."""
    per = -(-T // world)
    per = -(-per // align) * align
    start, end = min(T, rank * per), min(T, (rank + 1) * per)
    return {"start": start, "end": end, "lo": max(0, start - halo_left), "hi": min(T, end + halo_right),
            "halo_left": halo_left, "halo_right": halo_right}


def exchange_halo(x_shard, plan, rank, world, group=None):
    """x_shard holds frames [start, end) of a (B, C, T) tensor split along time.  Fetch the missing
    [lo, start) from the left neighbour and [end, hi) from the right one; returns the extended tensor for frames
    [lo, hi).  All sends and receives of a rank go out as ONE batch (ncclGroupStart/End under NCCL: neighbour
    exchanges posted one by one would deadlock on a single stream).  Halos must not span more than one neighbour."""
    import torch.distributed as dist
    need_l, need_r = plan["start"] - plan["lo"], plan["hi"] - plan["end"]
    hl, hr = plan["halo_left"], plan["halo_right"]
    B, C, n = x_shard.shape
    assert world == 1 or (hl <= n and hr <= n), "halo wider than a shard"
    left = x_shard.new_empty((B, C, need_l))
    right = x_shard.new_empty((B, C, need_r))
    ops = []
    if rank > 0 and need_l > 0:
        ops.append(dist.P2POp(dist.irecv, left, rank - 1, group))
    if rank < world - 1 and need_r > 0:
        ops.append(dist.P2POp(dist.irecv, right, rank + 1, group))
    # what the neighbours need from me mirrors what I need from them (same halo widths everywhere)
    if rank < world - 1 and hl > 0:
        ops.append(dist.P2POp(dist.isend, x_shard[:, :, -hl:].contiguous(), rank + 1, group))
    if rank > 0 and hr > 0:
        ops.append(dist.P2POp(dist.isend, x_shard[:, :, :hr].contiguous(), rank - 1, group))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    return torch.cat([left, x_shard, right], 2)


class PeerHaloExchange(object):
    """The same exchange through NVLink peer memory instead of NCCL point-to-point operations.  Every rank owns one
    symmetric buffer (torch symmetric memory: the allocation of each process of the node is mapped into all the
    others); a rank WRITES the frames its neighbours miss straight into THEIR buffers -- peer stores over NVLink /
    NVSwitch issued by a copy kernel on the current stream --, raises a signal in the neighbour's signal pad and waits
    for its own two signals.  No NCCL on the data path, no staging copy, nothing on the host.

    Buffer of a rank: 2 slots (step parity) x [left halo (B, C, halo_left) | right halo (B, C, halo_right)].  A neighbour
    may start writing step k+1 while this rank still reads step k (other slot).  Before it rewrites the slot of step k at
    step k+2 it waits for this rank's ACK of step k (a signal raised after the read, stream-ordered): with one-directional
    halos (any causal network: halo_right = 0) the writer never waits on the reader otherwise, and nothing else would
    keep it from running two steps ahead.  Channels: 2 * parity + side for the data, 4 + 2 * parity + side for the acks."""

    def __init__(self, B, C, halo_left, halo_right, dtype, rank, world, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.B, self.C, self.hl, self.hr, self.dtype = B, C, int(halo_left), int(halo_right), dtype
        self.rank, self.world = rank, world
        self.slot_elems = B * C * (self.hl + self.hr)
        self.buf = symm.empty(max(1, 2 * self.slot_elems), dtype=dtype, device=torch.device("cuda", torch.cuda.current_device()))
        self.hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.step = 0

    def _slot(self, peer, side, parity):
        """(B, C, h) view of `peer`'s left (side 0) or right (side 1) halo slot of the given step parity."""
        h = self.hl if side == 0 else self.hr
        off = parity * self.slot_elems + (0 if side == 0 else self.B * self.C * self.hl)
        return self.hdl.get_buffer(peer, (self.B, self.C, h), self.dtype, off)

    def exchange(self, x_shard, plan):
        """x_shard: frames [start, end) of a (B, C, T) tensor -> frames [lo, hi) (same result as exchange_halo)."""
        rank, world, hl, hr = self.rank, self.world, self.hl, self.hr
        B, C, n = x_shard.shape
        assert (B, C) == (self.B, self.C) and x_shard.dtype == self.dtype and hl <= n and hr <= n
        par = self.step & 1
        reuse = self.step >= 2                                # this parity's slots were written two steps ago
        self.step += 1
        if rank < world - 1 and hl > 0:                       # my last frames -> right neighbour's LEFT halo slot
            if reuse:
                self.hdl.wait_signal(rank + 1, channel=4 + 2 * par)       # ... and it has read them
            self._slot(rank + 1, 0, par).copy_(x_shard[:, :, n - hl:])
            self.hdl.put_signal(rank + 1, channel=2 * par)
        if rank > 0 and hr > 0:                               # my first frames -> left neighbour's RIGHT halo slot
            if reuse:
                self.hdl.wait_signal(rank - 1, channel=4 + 2 * par + 1)
            self._slot(rank - 1, 1, par).copy_(x_shard[:, :, :hr])
            self.hdl.put_signal(rank - 1, channel=2 * par + 1)
        need_l, need_r = plan["start"] - plan["lo"], plan["hi"] - plan["end"]
        parts = []
        if rank > 0 and hl > 0:
            self.hdl.wait_signal(rank - 1, channel=2 * par)
            parts.append(self._slot(rank, 0, par)[:, :, hl - need_l:])
        parts.append(x_shard)
        if rank < world - 1 and hr > 0:
            self.hdl.wait_signal(rank + 1, channel=2 * par + 1)
            parts.append(self._slot(rank, 1, par)[:, :, :need_r])
        out = torch.cat(parts, 2) if len(parts) > 1 else x_shard
        # acks: the halo slots of this step have been read (the cat above is ordered before these on the stream)
        if rank > 0 and hl > 0:
            self.hdl.put_signal(rank - 1, channel=4 + 2 * par)
        if rank < world - 1 and hr > 0:
            self.hdl.put_signal(rank + 1, channel=4 + 2 * par + 1)
        return out


def time_sharded_forward(forward_fn, x_ext, plan, T, out_extra=0):
    """Run `forward_fn` on the halo-extended input of one rank and keep the output frames this rank owns.
    Output frame g of the full read corresponds to local frame g - lo.  `out_extra` = frames the network appends
    (RawCTCNet: feature_kwidth - 1, emitted by the last rank only)."""
    y = forward_fn(x_ext)
    lo, start, end = plan["lo"], plan["start"], plan["end"]
    stop = end + (out_extra if end == T else 0)
    return y[:, :, start - lo:stop - lo]
