"""rlsim-b200: B200-native library-construction hot path of rlsim (priming/fragmentation, PCR,
size selection, final draw) behind the reference's Pooler / Sampler interfaces.

Host-side mirror (Python, over the C ABI in include/rlsim_gpu.h):
    args      - CLI flags                      (/root/reference/src/args.go)
    mixture   - -d / -a mixture mini-language  (target_mix.go)
    fasta     - FASTA input                    (fasta.go, input.go)
    report    - JSON report                    (report.go, fragstats.go)
    pool      - GpuPool  = Pooler.InitTranscripts   (pool.go:83-135)
    sampler   - GpuSampler = Sampler.SampleFragments (sampler.go:108-154)
    main      - driver, `python -m rlsim_b200` (main.go)
The compute lives in rlsim_b200/csrc/*.cu -> rlsim_b200/librlsim_gpu.so.  There is no CPU fallback.
"""
__version__ = "0.1.0"
