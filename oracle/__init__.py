"""CPU ORACLE for the MAS hot path -- TEST INFRASTRUCTURE, not the product.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.  art_tts_b200/ never imports it and has no CPU fallback.

Parity status: PINNED against the reference's own compiled Cython kernel
(oracle/_ref, built by oracle/build_ref.py from /root/reference) through the golden
vectors in tests/golden/ (made by tests/golden/make_golden.py in the build container).

Contents
  mas_oracle.c        plain-C restatement of core.pyx:9-45 and the tts.py:483-505 block
  mas_oracle.py       numpy/ctypes front-end mirroring monotonic_align/__init__.py:8-23
  build_ref.py        compiles the reference's core.pyx into oracle/_ref/ (git-ignored)
  ref.py              loader for oracle/_ref (the real reference kernel, when built)
"""
from .mas_oracle import (  # noqa: F401
    align_mu_y,
    align_mu_y_grad,
    build_oracle,
    crop_offsets,
    crop_segments,
    duration_loss,
    duration_targets,
    durations_from_path,
    frame_index,
    generate_path,
    inference_alignment,
    log_prior,
    log_prior_f64,
    maximum_path,
    maximum_path_c,
    maximum_path_rowsweep,
    oracle_threads,
    prior_loss,
    sequence_mask,
)
