"""CPU oracle for the ReconVAT hot path (Mel front-end + VAT perturbation loop).

TEST INFRASTRUCTURE ONLY.  Nothing under ``reconvat_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
or as the timed CPU baseline -- never as the product path.

Parity status: PINNED against the reference's own Python, executed in the
build container through ``oracle/reference_loader.py`` (the reference ships no
tests or golden vectors of its own -- SURVEY.md section 8c).  The outputs of that
run are committed under ``tests/golden/`` together with the generating script
``oracle/make_golden.py``; ``tests/test_oracle_golden.py`` checks this
restatement against them.

Files
-----
nnaudio_restate.py   restatement of the three un-vendored nnAudio==0.2.0 /
                     librosa 0.7 helpers the reference front-end calls
frontend.py          reflect pad -> STFT contraction -> power -> Mel -> log ->
                     imagewise min-max -> transpose   (model/Spectrogram.py,
                     model/utils.py, model/self_attention_VAT.py:1100-1104)
vat.py               the VAT perturbation loop, closed form and autograd form
                     (model/self_attention_VAT.py:101-255 and siblings)
attention.py         MutliHeadAttention1D.forward (model/self_attention_VAT.py:61-88)
decoding.py          extract_notes_wo_velocity / notes_to_frames (model/decoding.py)
cpu_path.py          the whole step as the reference's ATen op sequence: the timed
                     CPU baseline of bench.py
reference_loader.py  imports the *unmodified* reference modules (build
                     container only; /root/reference is absent on the GPU box)
make_golden.py       regenerates tests/golden/*.npz from the reference
"""
