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
* ``mel_ref``      — the reference's own ``audio.py`` / ``hparams.py`` and
  ``MelASR.run_step`` are executed in the build container and pin everything
  around the two librosa calls (``tests/golden/mel_chain_golden.npz``,
  ``mel_window_golden.npz``); librosa itself is absent (unpinned in
  requirements.txt): ``librosa.stft`` / ``librosa.filters.mel`` are restated
  from the published algorithm and cross-checked against ``torch.stft`` and
  ``torchaudio.functional.melscale_fbanks`` — **parity unpinned** w.r.t. a
  real librosa install for those two functions only.
* ``musetalk_ref`` — positional encoding and the VAE pre/post-processing are
  pinned to the reference's own modules (``pe_golden.npz``,
  ``vae_glue_golden.npz``); the diffusers UNet / VAE arithmetic is **parity
  unpinned** (package and checkpoints absent everywhere).
* the wav2lip glue (batch assembly, x255, truncation, paste-back) is pinned to
  the reference's own ``LipReal.inference_batch`` / ``paste_back_frame``
  (``lipreal_golden.npz``); the Whisper window indices to its
  ``BaseASR._get_sliced_feature`` (``slice_golden.npz``).
* ``yuv_ref``      — pinned bit-exact against OpenCV; **parity unpinned**
  against libswscale (absent), which the reference's encoder path uses.
"""
