"""The reference's train.py:9 does `from config import *`, but ships no config module.
This empty stand-in lets that star-import succeed."""
