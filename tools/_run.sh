O=gpurun_out
python -m pytest tests -x -q -m gpu -k "multichannel" 2>&1 | tail -2
python tools/decode_shapes.py 4:2048:512000 > $O/shapes.txt 2>&1
cat $O/shapes.txt
