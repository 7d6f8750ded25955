"""CPU oracle for the hot path (TEST INFRASTRUCTURE ONLY).

This package restates, in plain PyTorch fp32 on the CPU, the arithmetic of the reference's
BSVD -> RealESRGAN path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; the product package
(``sharkshark-4k_b200``) never does and fails loudly when its CUDA library is missing.

Pinning status:
  * SRVGGNetCompact, BSVD: restatement checked against the reference's own modules imported
    from /root/reference (oracle/reference_import.py, tests/golden/make_golden.py) -- max |diff| 0
    on the committed fixtures.
  * RRDBNet: lives in the un-vendored pip package ``basicsr`` (no version pinned by the reference;
    reached through ``realesrgan`` @ 5ca1078535923d485892caee7d7804380bfc87fd, README.md:62).
    Restated from the published architecture; pinned only by parameter counts and state-dict key
    names -> **parity unpinned** for the RRDBNet arithmetic itself.
  * the reference holds no golden vectors or known-answer tests for this path.
"""
