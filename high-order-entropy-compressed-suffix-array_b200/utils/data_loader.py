"""Drop-in for the reference's utils/data_loader.py:3-7 (gzip -> latin-1 text)."""
import gzip


def load_text(path, size_limit=None):
    with gzip.open(path, 'rt', encoding='latin-1') as f:
        if size_limit:
            return f.read(size_limit)
        return f.read()
