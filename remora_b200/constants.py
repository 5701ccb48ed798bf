"""Default values shared with the reference (src/remora/constants.py:1-9); values only."""
DEFAULT_NN_SIZE = 64
DEFAULT_BATCH_SIZE = 2048
DEFAULT_CHUNK_CONTEXT = (200, 200)
DEFAULT_KMER_CONTEXT_BASES = (4, 4)
DEFAULT_KMER_LEN = sum(DEFAULT_KMER_CONTEXT_BASES) + 1
