timeout 600 python -m pytest tests/test_sql_gpu.py tests/test_sql_tc_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 120 python tools/time_sql.py 2>&1 | head -6
timeout 120 python tools/time_sql.py 8 160 512 128 128 2>&1 | head -6
