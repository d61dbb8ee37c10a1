set -x
SEL='tests/test_nvbio_tb_gpu.py::test_golden_vectors_through_the_c_abi tests/test_nvbio_tb_gpu.py::test_traceback_in_several_chunks tests/test_nvbio_tb_gpu.py::test_banded_short_texts_indexed_pairs_and_ascii tests/test_wave_gpu.py::test_wave16_semiglobal_ragged tests/test_banded_tb_gpu.py tests/test_mixed_dispatch_gpu.py'
timeout 600 python -m pytest $SEL -m gpu -q -x 2>&1 | tail -3 > gpurun_out/r2_san_plain.log; cat gpurun_out/r2_san_plain.log
grep -q passed gpurun_out/r2_san_plain.log && ! grep -q failed gpurun_out/r2_san_plain.log && \
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2_sanitizer_memcheck.log python -m pytest $SEL -m gpu -q -x > gpurun_out/r2_san_run.log 2>&1
echo sanitizer rc=$?; tail -3 gpurun_out/r2_san_run.log; tail -8 gpurun_out/r2_sanitizer_memcheck.log
