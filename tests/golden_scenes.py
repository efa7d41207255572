"""Loads tests/golden/make_golden.py (scene table + helpers) and digests.json for the tests."""
import importlib.util
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)

SCENES = make_golden.SCENES
SMALL = make_golden.SMALL
DIGESTS = json.load(open(os.path.join(HERE, "golden", "digests.json")))["scenes"]
digest = make_golden.digest
