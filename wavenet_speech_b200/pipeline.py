"""Host-buffer pipeline: run a drop-in module over a batch that lives in pinned HOST memory, overlapping the
host->device copy of chunk i+1, the kernels of chunk i and the device->host copy of chunk i-1 on three CUDA
streams.  Reads in a batch are independent through every op on the path, so splitting the batch is exact.

The reference's callers do `model(signal.cuda())` followed by `.cpu()` (legacy_code/run_raw_ctc.py:57-59,
pretrain_tnt.py:145,159): the copies and the compute are serialised on one stream.  PCIe is full duplex and the
copy engines are independent of the SMs, so with >= 3 chunks the step costs max(copy-in, compute, copy-out)
instead of their sum."""
import torch


class HostPipeline(object):
    """`fn` (default: the model itself) is what runs on each device chunk; `HostPipeline(net, fn=net.forward_levels)`
    streams (B, T) uint8 / integer level tensors instead of (B, 256, T) one-hot tensors: 1/512 of the bytes."""

    def __init__(self, model, chunks=4, device=None, fn=None, graph=False):
        """graph=True: the forward of a (full-size) chunk is captured once per staging slot as a CUDA graph and replayed --
        one launch per chunk instead of ~25 Python-issued ones, so the host is far ahead of the GPU from the first step
        after a synchronisation (a 20-step timed region no longer depends on the host thread never being descheduled)."""
        self.model = model if fn is None else fn
        self.graph = bool(graph)
        self._graphs = {}          # staging slot -> (GraphedForward on that slot's buffer, event: its output was copied out)
        self._graph_key = None     # what the captured graphs depend on besides the buffers: weights, activation format
        self._module = model if isinstance(model, torch.nn.Module) else getattr(self.model, "__self__", None)
        self.chunks = int(chunks)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.s_in = torch.cuda.Stream(self.device)
        self.s_out = torch.cuda.Stream(self.device)
        self._bufs = None
        # per staging slot: the event after which the slot may be overwritten (its last consumer's kernels are done).
        # Instance state, not per-submit state: the slots persist across submits, so the first copy-ins of batch k+1
        # must wait for the last chunks of batch k that still read them.
        self._buf_free = [None, None]

    def _buffers(self, n, x_shape, x_dtype):
        key = (n, tuple(x_shape), x_dtype)
        if self._bufs is None or self._bufs[0] != key:
            xs = [torch.empty(x_shape, dtype=x_dtype, device=self.device) for _ in range(min(n, 2))]
            self._bufs = (key, xs)
            self._graphs = {}                    # graphs were captured on the old buffers
            # fresh blocks come from the compute stream's pool: whatever used that memory before is ordered on the
            # compute stream, so the copy stream waits for this point before its first write
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            self._buf_free = [ev, ev]
        return self._bufs[1]

    def _capture_key(self):
        """A captured graph holds the weight packs and the activation format of its capture: new weights (optimiser
        step, load_state_dict) or another `tc_precision` / `reduced_precision` setting mean a new capture."""
        from . import fastpath
        mod = self._module
        wkey = None if not isinstance(mod, torch.nn.Module) else tuple((p.data_ptr(), p._version) for p in mod.parameters())
        return (wkey, fastpath._PRECISE, fastpath._REDUCED, fastpath.HI_ONLY_TAIL)

    def __call__(self, x_host, y_host=None):
        """x_host: pinned CPU tensor (B, C, T).  Returns y_host (pinned CPU tensor), valid when the call returns
        (it synchronises on the last copy-out)."""
        y_host = self.submit(x_host, y_host)
        self.wait()
        return y_host

    def wait(self):
        """Block until every copy-out submitted so far has landed in host memory."""
        self.s_out.synchronize()

    @torch.no_grad()
    def submit(self, x_host, y_host=None):
        """Enqueue one batch without waiting for it: back-to-back submits overlap the tail of one batch (last
        kernels, last copy-out) with the head of the next (first copy-in).  `y_host` is valid after wait()."""
        assert not x_host.is_cuda and x_host.is_pinned(), "HostPipeline expects a pinned host tensor"
        B = x_host.shape[0]
        n = max(1, min(self.chunks, B))
        bounds = [(B * i // n, B * (i + 1) // n) for i in range(n)]
        per = max(e - s for s, e in bounds)
        xbuf = self._buffers(n, (per,) + tuple(x_host.shape[1:]), x_host.dtype)
        cur = torch.cuda.current_stream(self.device)
        if self.graph:
            key = self._capture_key()
            if key != self._graph_key:
                self._graphs, self._graph_key = {}, key
        in_done = [torch.cuda.Event() for _ in range(n)]
        comp_done = [torch.cuda.Event() for _ in range(n)]
        buf_free = self._buf_free
        outs = []
        for i, (s, e) in enumerate(bounds):
            slot = i % len(xbuf)
            with torch.cuda.stream(self.s_in):
                if buf_free[slot] is not None:
                    self.s_in.wait_event(buf_free[slot])          # kernels of chunk i-2 have consumed the slot
                xd = xbuf[slot][:e - s]
                xd.copy_(x_host[s:e], non_blocking=True)
                in_done[i].record(self.s_in)
            cur.wait_event(in_done[i])
            graphed = self.graph and (e - s) == per
            if graphed:
                g = self._graphs.get(slot)
                if g is None:                    # first use of this slot: warm up + capture on the slot's own buffer
                    g = [GraphedForward(self.model, xbuf[slot], static_input=True), None]
                    self._graphs[slot] = g
                if g[1] is not None:
                    cur.wait_event(g[1])         # the graph's output buffer: its previous content has been copied out
                g[0].graph.replay()
                y = g[0].y
            else:
                y = self.model(xd)
            comp_done[i].record(cur)
            buf_free[slot] = comp_done[i]
            if y_host is None:
                y_host = torch.empty((B,) + tuple(y.shape[1:]), dtype=y.dtype).pin_memory()
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(comp_done[i])
                y_host[s:e].copy_(y, non_blocking=True)
                if graphed:
                    g[1] = torch.cuda.Event()
                    g[1].record(self.s_out)
                else:
                    y.record_stream(self.s_out)
            outs.append(y)
        return y_host


class GraphedForward(object):
    """model(x) for ONE input shape captured into a CUDA graph: a forward of a small batch is ~20 launches of a few
    microseconds each, so issuing them from Python (ctypes call + tensor-map encode per launch) costs more than running
    them; the graph replays the whole sequence with one launch.  Every kernel of the path takes caller-owned buffers
    and a stream and never synchronises, so the capture needs no special casing: the tensor maps encoded at capture
    time point at the graph's private buffers.  `g(x)` copies x into the static input, replays, and returns the static
    output tensor (valid until the next call)."""

    def __init__(self, model, example, warmup=2, static_input=False):
        """static_input=True: `example` itself is the graph's input buffer (the caller refills it and calls
        `self.graph.replay()`; the result is `self.y`)."""
        assert example.is_cuda, "GraphedForward needs a CUDA example input"
        self.model = model
        self.x = example if static_input else example.detach().clone()
        side = torch.cuda.Stream(example.device)
        side.wait_stream(torch.cuda.current_stream(example.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                       # weight packs, function attributes, allocator warm-up
                model(self.x)
        torch.cuda.current_stream(example.device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.y = model(self.x)

    def __call__(self, x):
        assert x.shape == self.x.shape and x.dtype == self.x.dtype, "GraphedForward was captured for another shape"
        self.x.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.y
