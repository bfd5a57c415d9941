"""TEST INFRASTRUCTURE -- the oracle for the OctFusion denoising U-Net hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import anything from here.  The product package `octfusion_b200` never does; it
fails loudly when its CUDA library is missing.

Contents
  ocnn_shim/      restated third-party `ocnn` symbols (parity unpinned at that boundary)
  ref_import.py   imports the *unmodified* reference modules from /root/reference (only in
                  the build container; used to pin `restate.py` and to generate tests/golden/)
  restate.py      torch-CPU restatement of every hot-path function, each citing the
                  reference file:line it follows (this is what travels to the GPU box)
  gen_golden.py   regenerates tests/golden/*.pt from the real reference
"""
