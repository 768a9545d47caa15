O=gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
QOA_B200_TMAP=1 python -m pytest tests -x -q -m gpu -k "multichannel or mixed or shuffled" 2>&1 | tail -2
python tools/decode_shapes.py 3:2048:512000 4:2048:512000 5:1024:512000 6:1024:512000 7:1024:512000 8:1024:512000 > $O/shapes.txt 2>&1
cat $O/shapes.txt
