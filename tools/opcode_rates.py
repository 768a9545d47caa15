"""Per-opcode integer throughput on the current GPU (development aid for the kernel instruction diet)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import qoaudio_b200 as q
eng = q.Engine(0)
names = ["IMAD", "LOP3", "IMAD+LOP3 1:1", "I2IP.SAT", "VIMNMX", "SHF", "PRMT", "LEA.HI.SX32", "IMAD.HI", "IADD3/IMAD.IADD",
         "ISETP+SEL", "IMAD.WIDE", "IMAD imm", "IMUL (R,R,R,RZ)", "LOP3 imm", "SHF 1-reg", "OR imm", "IADD3 3-reg",
         "IMAD+SHF1 1:1", "IMAD:LOP3:SHF1 2:1:1", "5 IMAD:4 FFMA:8 LOP3", "9 IMAD:8 LOP3"]
if len(sys.argv) > 1:
    for k in sys.argv[1:]:
        print(f"{int(k):2d} {names[int(k)]:22s} {eng.int_issue_peak(int(k)) / 1e12:7.2f} T thread-ops/s")
    sys.exit(0)
for k, n in enumerate(names):
    print(f"{k:2d} {n:18s} {eng.int_issue_peak(k) / 1e12:7.2f} T thread-ops/s")
