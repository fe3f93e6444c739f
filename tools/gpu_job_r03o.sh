python -m pytest tests -m gpu -q -x 2>&1 | grep -v Warning | tail -5
python tools/chol_probe.py | cut -c1-330
