timeout 100 python -m pytest tests/test_gpu_pipeline.py -x -q 2>&1 | tail -3
