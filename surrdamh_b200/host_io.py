"""Double-buffered host I/O for a block of lockstep chains.

What a host-side user of the chains exchanges with the GPU per block of K steps: the chains' current samples go up
(pinned host memory -> HBM), the block's product comes back -- final samples, counters, the running first / second
moments of every chain and the accept / reject flag of every step (the reference's product is the weighted chain
written by its samplers, classes_SAMPLER.py:130-134; flags + proposals reconstruct it, `trace_csv.py`).  The transfers
run on a copy stream of their own, beside the chain kernel:

    upload(b+1)    during block b:  pinned u_in -> a device staging buffer             (copy stream)
    begin(b+1)     at its start:    staging -> blk.u, D2D                              (sampling stream, waits for the upload)
    finish(b)      at its end:      blk.u / counters / moments -> device staging, D2D  (sampling stream), then staging and
                                    the block's flag trace -> pinned set b mod 2       (copy stream)
    result(b)      host:            waits for block b's download, returns the pinned set; a set is reused two blocks later

so the PCIe time of block b overlaps the kernel of block b+1, and the host never blocks on the block it has just issued.
"""
import torch


class BlockIO:
    def __init__(self, blk, K, flags=True):
        self.blk, self.K = blk, int(K)
        dev, d, C_ = blk.device, blk.d, blk.C
        if blk.moments is None:
            blk.moments = torch.zeros((blk.n_mom, C_), dtype=torch.float64, device=dev)
        self.copy = torch.cuda.Stream(device=dev)
        self.u_in = torch.empty((d, C_), dtype=torch.float64).pin_memory()             # the caller fills this
        self.u_stage = [torch.empty((d, C_), dtype=torch.float64, device=dev) for _ in range(2)]
        self.up_done = [None, None]
        self.sets, self.stage, self.down_done = [], [], [None, None]
        for _ in range(2):
            self.sets.append({"u": torch.empty((d, C_), dtype=torch.float64).pin_memory(),
                              "counters": torch.empty((4, C_), dtype=torch.int32).pin_memory(),
                              "moments": torch.empty((blk.n_mom, C_), dtype=torch.float64).pin_memory(),
                              "flags": torch.empty((self.K, C_), dtype=torch.uint8).pin_memory() if flags else None})
            self.stage.append({"u": torch.empty_like(blk.u), "counters": torch.empty_like(blk.counters),
                               "moments": torch.empty_like(blk.moments),
                               "flags": torch.empty((self.K, C_), dtype=torch.uint8, device=dev) if flags else None})
        self.h2d_bytes = self.u_in.numel() * 8
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.sets[0].values() if t is not None)

    def upload(self, b):
        """Start the upload of `u_in` for block b (call any time before begin(b); the caller must not touch u_in until
        begin(b) has been called)."""
        with torch.cuda.stream(self.copy):
            self.u_stage[b & 1].copy_(self.u_in, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        self.up_done[b & 1] = ev

    def begin(self, b):
        """The uploaded samples become the chains' current samples (sampling stream); follow with blk.prepare()."""
        if self.up_done[b & 1] is None:
            self.upload(b)
        torch.cuda.current_stream().wait_event(self.up_done[b & 1])
        self.blk.u.copy_(self.u_stage[b & 1])
        self.up_done[b & 1] = None

    def finish(self, b, trace=None):
        """Queue the download of block b's product (after blk.run on the sampling stream)."""
        blk, st, hs = self.blk, self.stage[b & 1], self.sets[b & 1]
        if self.down_done[b & 1] is not None:
            self.down_done[b & 1].synchronize()            # the set's previous content (block b-2) has been delivered
        st["u"].copy_(blk.u)
        st["counters"].copy_(blk.counters)
        st["moments"].copy_(blk.moments)
        with_flags = trace is not None and hs["flags"] is not None
        if with_flags:
            st["flags"].copy_(trace.flags[:self.K])         # the trace buffer is reused by a later block: stage it too
        ready = torch.cuda.Event()
        ready.record()
        with torch.cuda.stream(self.copy):
            self.copy.wait_event(ready)
            hs["u"].copy_(st["u"], non_blocking=True)
            hs["counters"].copy_(st["counters"], non_blocking=True)
            hs["moments"].copy_(st["moments"], non_blocking=True)
            if with_flags:
                hs["flags"].copy_(st["flags"], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        self.down_done[b & 1] = ev
        return ev

    def result(self, b):
        """Host: block b's product (waits for its download)."""
        if self.down_done[b & 1] is not None:
            self.down_done[b & 1].synchronize()
        return self.sets[b & 1]

    def drain(self):
        for ev in self.down_done:
            if ev is not None:
                ev.synchronize()
