mkdir -p gpurun_out/D
O=gpurun_out/D
timeout 600 python -m pytest tests/test_next_rows.py -m gpu -q -x -k "build_trt and redconv and int8" 2>&1 | tail -40 > $O/t1.log
timeout 600 python -m pytest tests/test_next_rows.py -m gpu -q -x -k "device_jpeg" 2>&1 | tail -40 > $O/t2.log
cat $O/t1.log $O/t2.log | cut -c1-250
