"""ncu target: C4 banded + CIGAR on the s16x2 lanes (k_banded_tb_s16 + k_get_tb), 65536 pairs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
b4 = synth.config4(n, seed=4)
mq = int(b4.q_lens.max()) + 7; mt = int(b4.t_lens.max()) + 7
with Aligner(max_query_len=mq, max_target_len=mt, max_n_alns=b4.n, n_streams=1, algo=capi.BANDED, k_band=64, start=capi.WITH_TB) as al:
    for _ in range(2):
        al.align(b4); print(al.kernel_name(0), al.last_ms(0), flush=True)
