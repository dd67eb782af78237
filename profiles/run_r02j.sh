timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q --tb=line -k "adam" 2>&1 | tail -12
