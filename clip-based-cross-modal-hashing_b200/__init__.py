"""cmh_b200 - B200-native retrieval evaluation (Hamming / mAP@K / precision@N / PR curve / top-K) behind the
call signatures of the reference's `utils/calc_utils.py`.  Import as ``cmh_b200`` (see ../cmh_b200/__init__.py).
"""
__version__ = "0.1.0"
