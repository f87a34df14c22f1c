"""Readers for the pickle stream written by OutputManager (reference pyds/user_utils.py)."""
import pickle


def read_output_file(filepath):
    objs = []
    with open(filepath, "rb") as f:
        while True:
            try:
                objs.append(pickle.load(f))
            except EOFError:
                break
    return objs


def read_last_file(filepath):
    last = None
    with open(filepath, "rb") as f:
        while True:
            try:
                last = pickle.load(f)
            except EOFError:
                break
    return last
