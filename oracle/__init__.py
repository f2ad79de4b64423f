"""CPU oracle for the dysfluency audio front-end  --  TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy restatement of the arithmetic the reference
(kishormb/Recognizing-Speech-Dysfluencies-in-Stuttering) inherits from librosa,
noisereduce and soundfile on its hot path:

    pipeline1.py:126-146   clean_audio_and_cache   (denoise -> normalise -> PCM-16)
    pipeline1.py:206-265   extract_audio_features / extract_features (149-vector)
    pipeline1.py:429-440   cached_extract_features (.npy cache)
    pipeline1.py:470-473   StandardScaler fit/transform ("global CMVN")

It is the *checker*.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.
The shipped product (``recognizing-speech-dysfluencies-in-stuttering_b200``)
never imports, links or falls back to anything in here.

Parity status
-------------
* feature function (a6-a13)  : PINNED  -- reproduces the 888 committed
  ``clear_audio/*.wav -> cache_features/*_clean_feats.npy`` pairs of the
  reference within atol 1e-3 / rtol 1e-4 (tests/test_oracle_golden.py).
* CMVN (StandardScaler)      : PINNED  -- reproduces output_results/scaler_after.pkl.
* normalise + PCM-16 quantise: pinned by property (every committed WAV peaks at full scale).
* denoise (noisereduce)      : PINNED STATISTICALLY -- noisereduce is not installable in
  the build container; oracle/denoise.py restates its non-stationary spectral gate from
  the published algorithm (same scipy routines: filtfilt, fftconvolve) and reproduces the
  reference's 888 mp3 -> clear_audio/*.wav pairs to a median 70.9 dB, with every perturbed
  default 25 - 50 dB worse on every clip (tests/test_denoise_pin.py).
* load (librosa.load)        : lengths PINNED exactly (888/888, libmpg123's gapless trimming +
  ceil), samples statistically: oracle/resample.py restates soxr's published HQ recipe
  (soxr itself stays unpinned); MP3 frames are decoded by the FFmpeg libavcodec that
  ships in the image (recognizing-speech-dysfluencies-in-stuttering_b200/mp3io.py).
"""
