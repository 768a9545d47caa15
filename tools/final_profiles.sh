#!/bin/bash
# Round-end evidence run (one B200): bench lines, ncu launch list, ncu full captures, compute-sanitizer.
# Everything lands in gpurun_out/; tools/summarize_ncu.py + a copy into profiles/ happen afterwards on the CPU side.
set -x
O=gpurun_out
python bench.py > $O/final_bench_n1.json 2> $O/final_bench_n1.err
python bench.py --impl reference > $O/final_bench_ref.json 2> $O/final_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/final_launches.csv \
    python bench.py --steps 3 --warmup 3 --streams 2048 --enc-streams 2048 --no-cpu --e2e-streams 32 --no-single > $O/final_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_fast_kernel -s 3 -c 1 -o $O/final_dec -f \
    python bench.py --steps 1 --warmup 3 --streams 1024 --no-encode --no-cpu --no-e2e --no-single > $O/final_dec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 3 -c 1 -o $O/final_enc -f \
    python bench.py --steps 1 --warmup 3 --streams 64 --enc-streams 4096 --enc-frames 4 --no-cpu --no-e2e --no-single > $O/final_enc.log 2>&1
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "mixed or stereo_batch or mono_batch or multichannel or variants or truncated or tensor_rows" > $O/final_memcheck.log 2>&1
compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "stereo_batch or mono_batch or multichannel_batch_fast_path and 6 or kat" > $O/final_racecheck.log 2>&1
tail -n 3 $O/final_memcheck.log $O/final_racecheck.log
tail -c 600 $O/final_bench_n1.json
