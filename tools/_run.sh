O=gpurun_out
rm -f $O/dv.txt
for v in "" "" ; do env $v python tools/dec_variants.py 8192 20 2>&1 | grep "^\[" >> $O/dv.txt; done
python tools/dec_variants.py 8192 10 1 2>&1 | grep "^\[" >> $O/dv.txt
cat $O/dv.txt
