"""Puts tools/ on sys.path so tests can import tools/configs.py (BASELINE.json configs 1-3)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
