"""Same job as the reference's ``utils/preprocessing/create-hdf5.py`` (which hard-codes its paths), with arguments:

    python utils/preprocessing/create-hdf5.py SRC DST
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from msrflute_b200.utils.preprocessing import main  # noqa: E402

if __name__ == "__main__":
    raise SystemExit(main(["tsv2hdf5"] + sys.argv[1:]))
