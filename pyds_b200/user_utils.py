"""Readers of the record stream that OutputManager appends to ``pyds_output.pkl`` -- the two helpers user scripts of the
reference import (``pyds/user_utils.py``: ``read_output_file``, ``read_last_file``), on top of one generator."""
import pickle


def iter_records(filepath):
    """The records of the file in the order they were written (one ``pickle.dump`` each)."""
    with open(filepath, "rb") as stream:
        size = stream.seek(0, 2)
        stream.seek(0)
        while stream.tell() < size:
            yield pickle.load(stream)


def read_output_file(filepath):
    """All records as a list."""
    return list(iter_records(filepath))


def read_last_file(filepath):
    """The record written last; None for an empty file."""
    last = None
    for last in iter_records(filepath):
        pass
    return last
