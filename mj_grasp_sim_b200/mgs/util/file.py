"""generate_unique_hash (/root/reference/mgs/util/file.py:21-30): scene directory names."""
import hashlib
import os
import time


def generate_unique_hash(length: int = 16) -> str:
    seed = f"{time.time()}-{os.getpid()}-{os.urandom(8).hex()}"
    return hashlib.sha256(seed.encode()).hexdigest()[:length]
