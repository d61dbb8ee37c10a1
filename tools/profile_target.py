"""Short, deterministic launch sequence for ncu (pack -> DP per mode on device-resident data)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from genomics_gpu_b200 import capi, synth
from genomics_gpu_b200.aligner import Aligner

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
batch = synth.config2(n, seed=2)
lib = capi.load()
for algo, start in ((capi.LOCAL, capi.WITH_START), (capi.SEMI_GLOBAL, capi.WITH_START), (capi.LOCAL, capi.WITHOUT_START), (capi.GLOBAL, capi.WITHOUT_START)):
    with Aligner(algo=algo, start=start, max_query_len=157, max_target_len=607, max_n_alns=n, n_streams=1) as al:
        al.align(batch)
        s = al.vec.a[0]
        for _ in range(reps):
            s.current_n_alns = n
            lib.gasalx_aln_device_resident(C.byref(s), batch.q_bytes.size, batch.t_bytes.size, n, C.byref(al.params))
            al.wait(0)
        print(algo, start, al.kernel_name(0), al.last_ms(0), flush=True)
