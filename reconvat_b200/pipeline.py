"""The hot-path step as one call: front-end + VAT on a batch of audio segments.

``HotPathStep(device)(audio)`` is what ``UNet.run_on_batch`` does before it reaches the network proper
(model/self_attention_VAT.py:1098-1106): Mel front-end, log, imagewise normalisation, transpose, then the
VAT loss against the given transcriber.  ``run_host`` feeds it from pinned host memory with the
host->device copy of batch i+1 overlapped with the kernels of batch i (copy stream + events).
"""
import torch

from . import Spectrogram, VAT

MEL_KW = dict(sr=16000, win_length=2048, n_mels=229, hop_length=512, fmin=30, fmax=8000,
              trainable_mel=False, trainable_STFT=False, verbose=False)   # model/self_attention_VAT.py:1027-1029


class HotPathStep:
    def __init__(self, model, device, xi=1e-6, eps=2.0, vat_cls=None):
        self.device = torch.device(device)
        self.model = model
        self.spectrogram = Spectrogram.MelSpectrogram(**MEL_KW).to(self.device)
        self.vat_loss = (vat_cls or VAT.UNet_VAT)(xi, eps, 1, False)
        self._copy_stream = None

    def __call__(self, audio):
        """audio: (B, L) float32 on the device.  Returns (vat_loss, r_norm_mean, spec, r_adv)."""
        spec = self.spectrogram.normalised_log_mel(audio)
        vat_loss, r_adv, r_norm = self.vat_loss(self.model, spec)
        return vat_loss, r_norm.abs().mean(), spec, r_adv

    def run_host(self, host_batches, results_host):
        """host_batches: iterable of pinned (B, L) float32 CPU tensors; results_host: pinned (n, 2) float32.
        Row i of results_host receives (vat_loss, r_norm_mean) of batch i.  Double-buffered: the copy of
        batch i+1 is issued on a side stream while batch i computes.  Returns the number of batches."""
        main = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        copy = self._copy_stream
        bufs, ready, freed = [None, None], [None, None], [None, None]
        it = iter(host_batches)

        def stage(slot, hb):
            if bufs[slot] is None or bufs[slot].shape != hb.shape:
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
            vat_loss, r_norm, _, _ = self(bufs[slot])
            ev = torch.cuda.Event()
            ev.record(main)
            freed[slot] = ev
            results_host[i].copy_(torch.stack((vat_loss.detach(), r_norm)), non_blocking=True)
            i += 1
            if nxt is None:
                return i
