python -m pytest tests -m gpu -x -q 2>&1 | tail -12
OUTLIER_FRAC=0 python tools/profile_filters.py | tail -1 | cut -c1-420
python tools/profile_filters.py | tail -1 | cut -c1-420
