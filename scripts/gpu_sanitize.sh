#!/bin/bash
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --print-limit 20 python scripts/sanitize.py > gpurun_out/r02_m_sanitizer_memcheck.log 2>&1; tail -n 25 gpurun_out/r02_m_sanitizer_memcheck.log
