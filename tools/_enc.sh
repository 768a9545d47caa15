QOA_B200_ENC_FAST=1 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
echo "--- fast"; STREAMS=1,592,1184 QOA_B200_ENC_FAST=1 python tools/enc_latency.py 5
echo "--- exact"; STREAMS=1,1184,4096 QOA_B200_ENC_FAST=0 python tools/enc_latency.py 5
