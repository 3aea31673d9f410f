mkdir -p gpurun_out/E
O=gpurun_out/E
timeout 600 python -m pytest tests/test_next_rows.py -m gpu -q 2>&1 | tail -8 > $O/t.log
cat $O/t.log | cut -c1-250
timeout 900 python tools/next_rows_time.py > $O/next_rows_time.jsonl 2> $O/next_rows_time.err
grep "8f-2" $O/next_rows_time.jsonl
tail -3 $O/next_rows_time.err
