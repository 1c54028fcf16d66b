"""The hot-path step as one call: front-end + VAT on a batch of audio segments.

``HotPathStep(model, device)(audio)`` is what ``UNet.run_on_batch`` does before it reaches the network proper
(model/self_attention_VAT.py:1098-1106): Mel front-end, log, imagewise normalisation, transpose, then the
VAT loss against the given transcriber.

Launch overhead: the step is ~10 short kernels (plus the caller's network), so eager launches leave the GPU idle
between them.  ``capture(buffers)`` records the whole step -- our kernels, ``torch.randn_like``, the network and
its backward -- into one CUDA graph per input buffer; ``replay(i)`` re-launches it with a single driver call.
``run_host`` feeds the step from pinned host memory with the host->device copy of batch i+1 overlapped with the
kernels of batch i (copy stream + events), replaying the graphs when they exist.
"""
import torch

from . import Spectrogram, VAT, _lib

MEL_KW = dict(sr=16000, win_length=2048, n_mels=229, hop_length=512, fmin=30, fmax=8000,
              trainable_mel=False, trainable_STFT=False, verbose=False)   # model/self_attention_VAT.py:1027-1029


class HotPathStep:
    def __init__(self, model, device, xi=1e-6, eps=2.0, vat_cls=None, precision=None):
        self.device = torch.device(device)
        self.model = model
        self.spectrogram = Spectrogram.MelSpectrogram(precision=precision, **MEL_KW).to(self.device)
        # strict=False: the step never synchronises; check() tests the NaN flags (eager and per graph) on demand
        self.vat_loss = (vat_cls or VAT.UNet_VAT)(xi, eps, 1, False, strict=False)
        # private reduction workspaces + the fused NaN flag / mean |d_hat| (VAT.Scratch); every captured graph gets
        # its own, because graphs replayed on different streams run the last-block reductions concurrently
        self._eager_scratch = VAT.Scratch(self.device, keep_d_hat=False)     # only mean |d_hat| leaves the step
        self._copy_stream = None
        self._graphs = []              # [(graph, input buffer, outputs, device flag)]
        self._lanes = []               # side streams of replay_many
        self.kernels_per_graph = 0

    def __call__(self, audio):
        """audio: (B, L) float32 on the device.  Returns (vat_loss, r_norm_mean, spec, r_adv)."""
        spec = self.spectrogram.normalised_log_mel(audio)
        self._n_rows = spec.numel() // spec.shape[-1]
        if self.vat_loss.scratch is None:
            self.vat_loss.scratch = self._eager_scratch
        vat_loss, r_adv, r_norm = self.vat_loss(self.model, spec)
        # r_norm.abs().mean() (model/self_attention_VAT.py:1149) comes out of the finalisation kernel
        r_norm_mean = self.vat_loss.last_r_norm_mean
        return vat_loss, (r_norm.abs().mean() if r_norm_mean is None else r_norm_mean), spec, r_adv

    # -- CUDA graphs ------------------------------------------------------------------------
    def capture(self, buffers, warmup=3):
        """Record one CUDA graph of the step per device input buffer (the graph reads that buffer in place; refill
        it with ``copy_`` between replays).  Every graph gets its own memory pool: with a shared pool the outputs
        of graph j alias intermediates of graph i and are clobbered when i replays (seen on the NaN flag)."""
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):                    # warm-up off the capture: lazy tables, cuBLAS handles, ...
            for i in range(warmup):
                self(buffers[i % len(buffers)])
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.vat_loss.check()
        self._graphs = []
        try:
            for buf in buffers:
                scratch = VAT.Scratch(self.device, keep_d_hat=False)
                scratch.stats(self._n_rows)               # sized before the capture (row count of the warm-up step)
                self.vat_loss.scratch = scratch
                g = torch.cuda.CUDAGraph()
                n0 = _lib.launch_count()
                with torch.cuda.graph(g):
                    out = self(buf)
                    out = (out[0].detach(),) + out[1:]
                self.kernels_per_graph = _lib.launch_count() - n0
                self._graphs.append((g, buf, out, self.vat_loss.last_flag, scratch))
        finally:
            self.vat_loss.scratch = self._eager_scratch
        return len(self._graphs)

    def replay(self, i):
        """Re-launch graph i on the current stream.  Returns (vat_loss, r_norm_mean, spec, r_adv): static tensors that
        the next replay of the same graph overwrites."""
        g, out = self._graphs[i][0], self._graphs[i][2]
        g.replay()
        return out

    def replay_many(self, order, streams=2):
        """Replay graphs ``order[0], order[1], ...`` with graph g always on stream ``g % streams``.  Consecutive steps
        are independent batches, so the HBM-bound VAT tail of one overlaps the tensor-bound contraction of the
        next (+18 % steps/s measured at B=32).  The current stream waits for all of them on return."""
        main = torch.cuda.current_stream(self.device)
        if len(self._lanes) < streams:
            self._lanes += [torch.cuda.Stream(self.device) for _ in range(streams - len(self._lanes))]
        lanes = self._lanes[:streams]
        for st in lanes:
            st.wait_stream(main)
        for g in order:
            with torch.cuda.stream(lanes[g % streams]):
                self._graphs[g][0].replay()
        for st in lanes:
            main.wait_stream(st)

    def check(self):
        """NaN/Inf assertion of the reference (model/self_attention_VAT.py:189-190) for the eager path and for every
        captured graph (synchronises)."""
        self.vat_loss.check()
        for entry in self._graphs:
            self.vat_loss.check(entry[3])

    # -- host-fed loop ----------------------------------------------------------------------
    def run_host(self, host_batches, results_host, use_graphs=True):
        """host_batches: iterable of pinned (B, L) float32 CPU tensors; results_host: pinned (n, 2) float32.
        Row i of results_host receives (vat_loss, r_norm_mean) of batch i.  Double-buffered: the copy of
        batch i+1 is issued on a side stream while batch i computes.  Returns the number of batches."""
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        copy = self._copy_stream
        graphs = use_graphs and len(self._graphs) >= 2
        bufs = [self._graphs[0][1], self._graphs[1][1]] if graphs else [None, None]
        ready, freed = [None, None], [None, None]
        it = iter(host_batches)

        def stage(slot, hb):
            if bufs[slot] is None or bufs[slot].shape != hb.shape:
                if graphs:
                    raise ValueError("run_host: batch shape %s does not match the captured buffers %s"
                                     % (tuple(hb.shape), tuple(bufs[slot].shape)))
                bufs[slot] = torch.empty(hb.shape, dtype=hb.dtype, device=self.device)
            with torch.cuda.stream(copy):
                if freed[slot] is not None:
                    copy.wait_event(freed[slot])          # do not overwrite a buffer the kernels still read
                bufs[slot].copy_(hb, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
                ready[slot] = ev

        nxt = next(it, None)
        if nxt is None:
            return 0
        stage(0, nxt)
        i = 0
        while True:
            slot = i & 1
            nxt = next(it, None)
            if nxt is not None:
                stage(slot ^ 1, nxt)
            main.wait_event(ready[slot])
            vat_loss, r_norm = (self.replay(slot) if graphs else self(bufs[slot]))[:2]
            results_host[i, 0].copy_(vat_loss.detach(), non_blocking=True)     # two 4-byte reads, no packing kernel
            results_host[i, 1].copy_(r_norm, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(main)
            freed[slot] = ev
            i += 1
            if nxt is None:
                return i
