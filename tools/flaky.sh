for i in 1 2 3 4 5; do timeout 300 python -m pytest tests/test_gpu_gpt.py -m gpu -q -k "teacher_forced_logits and tiny" 2>&1 | tail -1; done
