timeout 900 python -m pytest tests/test_mf_gpu.py tests/test_model_gpu.py -m gpu -q --tb=line -k "bloom or adam or mrr" 2>&1 | tail -8
python profiles/bench_bloom.py 2>&1 | tail -1
python profiles/bench_bloom.py --dense 2>&1 | tail -1
