"""Drop-in shim: ``from hansel import Hansel`` (gretel/gretel.py:7, gretel/util.py:4 of the
reference) resolves to the B200-native implementation when this repo is on sys.path."""
from gretel_b200.hansel import Hansel  # noqa: F401
