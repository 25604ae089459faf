"""Puts tests/golden (make_golden.py, the golden generator and case list) on sys.path."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
