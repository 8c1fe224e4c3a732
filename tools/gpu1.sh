set -x
nvidia-smi -L; nproc; free -g | head -2
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -30
