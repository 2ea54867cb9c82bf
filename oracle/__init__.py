"""CPU oracle for the lip-sync hot path — TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement of the reference's algorithm
(lipku/LiveTalking, paths cited per function).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / CPU
baseline — never as the thing shipped.  The product (``livetalking_b200``)
must never import this package.

Parity status: the reference ships no test, golden vector or
weights for this path ("parity unpinned" by its own tests).  We pin what can
be pinned:

* ``wav2lip_ref``  — pinned against the UNMODIFIED reference ``Wav2Lip``
  nn.Module imported from /root/reference in the build container
  (``tests/golden/make_golden.py`` → ``tests/golden/w2l_*.npz``).
* ``paste_ref``    — pinned bit-exact against ``cv2.resize`` (the library the
  reference calls) on randomized sizes, and against committed golden frames.
* ``mel_ref``      — librosa is absent (unpinned in requirements.txt): restated
  from librosa's published algorithm; cross-checked against ``torch.stft`` and
  ``torchaudio.functional.melscale_fbanks``.  **parity unpinned** w.r.t. a real
  librosa install.
"""
