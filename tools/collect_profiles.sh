#!/bin/bash
# CPU side of the evidence run: turn gpurun_out/final_* into the committed summaries under profiles/ (round tag $1).
R=${1:-r02}
O=gpurun_out
P=profiles
cp $O/final_bench_n1.json $P/${R}_bench_n1.json
cp $O/final_bench_ref.json $P/${R}_bench_reference_arm.json
cp $O/final_launches.csv $P/${R}_launches.csv
python tools/summarize_launches.py $O/final_launches.csv "python bench.py --steps 3 --warmup 3 --streams 2048 --enc-streams 2048 --no-cpu --e2e-streams 32 --no-single --no-extras" > $P/${R}_launches_summary.txt
python tools/summarize_ncu.py $O/final_dec.ncu-rep > $P/${R}_decode_fast_kernel_ncu_full.txt
python tools/summarize_ncu.py $O/final_enc.ncu-rep > $P/${R}_encode_kernel_ncu_full.txt
for c in 1 6 8; do python tools/summarize_ncu.py $O/final_dec_${c}ch.ncu-rep > $P/${R}_decode_${c}ch_ncu_full.txt; done
cp $O/final_decode_shapes.txt $P/${R}_decode_shapes.txt
cp $O/final_pipe_probe.txt $P/${R}_pipe_probe.txt
cp $O/final_core_probe.txt $P/${R}_core_probe.txt
for t in memcheck racecheck synccheck; do tail -n 12 $O/final_$t.log > $P/${R}_compute_sanitizer_$t.log; done
# SASS opcode histogram per kernel (built library, no GPU needed)
cuobjdump -sass qoaudio_b200/libqoa_b200.so | python tools/sass_histogram.py > $P/${R}_sass_opcode_histogram.txt
