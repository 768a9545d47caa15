O=gpurun_out
python bench.py --streams 1024 --enc-streams 512 --steps 3 > $O/b_small.json 2> $O/b_small.err || { tail -30 $O/b_small.err; exit 1; }
tail -c 3000 $O/b_small.json
( time python bench.py > $O/b_full.json 2> $O/b_full.err ) 2>&1 | tail -3
tail -5 $O/b_full.err
