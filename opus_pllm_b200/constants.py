"""Sequence constants of the reference (multi_modality_v1/constants.py:6-13)."""
IGNORE_INDEX = -100
DEFAULT_SEQ_TOKEN_INDEX = -200
DEFAULT_SEQ_TOKEN = "<seq>"
DEFAULT_SEQ_PATCH_TOKEN = "<seq_patch>"
DEFAULT_SEQ_START_TOKEN = "<seq_start>"
DEFAULT_SEQ_END_TOKEN = "<seq_end>"
SEQ_PLACEHOLDER = "<seq-placeholder>"
