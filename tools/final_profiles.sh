#!/bin/bash
# Round-end evidence run (one B200): bench lines, ncu launch list, ncu full captures, compute-sanitizer.
# Everything lands in gpurun_out/; tools/summarize_ncu.py + a copy into profiles/ happen afterwards on the CPU side
# (tools/collect_profiles.sh).
set -x
O=gpurun_out
make -s -C tools
python bench.py > $O/final_bench_n1.json 2> $O/final_bench_n1.err
python bench.py --impl reference > $O/final_bench_ref.json 2> $O/final_bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/final_launches.csv \
    python bench.py --steps 3 --warmup 3 --streams 2048 --enc-streams 2048 --no-cpu --e2e-streams 32 --no-single --no-extras > $O/final_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:decode_fast_kernel -s 3 -c 1 -o $O/final_dec -f \
    python bench.py --steps 1 --warmup 3 --streams 1280 --no-encode --no-cpu --no-e2e --no-single --no-extras > $O/final_dec.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:encode_kernel -s 3 -c 1 -o $O/final_enc -f \
    python bench.py --steps 1 --warmup 3 --streams 64 --enc-streams 4096 --enc-frames 4 --no-cpu --no-e2e --no-single --no-extras > $O/final_enc.log 2>&1
for c in 1 6 8; do
  ncu --set full --clock-control none -k regex:"decode_(multi|fast)_kernel" -s 1 -c 1 -o $O/final_dec_${c}ch -f \
      python tools/decode_shapes.py $c:512:512000 > $O/final_dec_${c}ch.log 2>&1
done
python tools/decode_shapes.py > $O/final_decode_shapes.txt 2>&1
./tools/pipe_probe > $O/final_pipe_probe.txt 2>&1
./tools/core_probe > $O/final_core_probe.txt 2>&1
K="mixed or stereo_batch or mono_batch or multichannel or variants or truncated or tensor_rows or multi_device"
compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$K" > $O/final_memcheck.log 2>&1
compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "stereo_batch or mono_batch or multichannel_batch_fast_path and 6 or kat" > $O/final_racecheck.log 2>&1
compute-sanitizer --tool synccheck python -m pytest tests/test_gpu_parity.py -m gpu -q -x \
    -k "stereo_batch or mono_batch or multichannel_batch_fast_path and 8 or kat or variants" > $O/final_synccheck.log 2>&1
tail -n 3 $O/final_memcheck.log $O/final_racecheck.log $O/final_synccheck.log
tail -c 400 $O/final_bench_n1.json
