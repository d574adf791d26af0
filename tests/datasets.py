"""Dataset fixtures: the reference's public g2o datasets (reference datasets/*.g2o), stored gzip'ed
under tests/golden/datasets so the GPU box (which has no /root/reference) can run them."""
import gzip
import os
import shutil
import tempfile

_HERE = os.path.dirname(os.path.abspath(__file__))
_CACHE = {}


def path(name):
    if name not in _CACHE:
        src = os.path.join(_HERE, "golden", "datasets", name + ".g2o.gz")
        d = tempfile.mkdtemp(prefix="spg_ds_")
        dst = os.path.join(d, name + ".g2o")
        with gzip.open(src, "rb") as f, open(dst, "wb") as o:
            shutil.copyfileobj(f, o)
        _CACHE[name] = dst
    return _CACHE[name]
