"""Append-only pickle log of every ``g`` evaluation -- the behaviour of the reference's
``pyds/output_manager.py`` (unchanged by the north star): the previous ``pyds_output.pkl`` is removed
at construction, a buffer ``{input, solver, simulator}`` is filled per inner iteration and appended."""
import pickle
from os import remove
from os.path import exists, join


class OutputManager:
    def __init__(self, folder):
        self.output_filename = "pyds_output"
        self.output_path_fnx = join(folder, self.output_filename + ".pkl")
        if exists(self.output_path_fnx):
            remove(self.output_path_fnx)
        self.clear_buffer()

    def clear_buffer(self):
        self.data = {"input": None, "solver": {}, "simulator": None}

    def add_solver_solution(self, container):
        self.data["solver"].update(container.copy())

    def add_simulator_solution(self, container):
        self.data["simulator"] = container.copy()

    def add_input(self, data):
        self.data["input"] = data.copy()

    def write_data_to_disk(self):
        with open(self.output_path_fnx, "ab+") as f:
            pickle.dump(self.data, f)

    def write_records_to_disk(self, records):
        """Append several records with one open(): the same byte stream as ``write_data_to_disk`` once per
        record (the two-stage manager logs one record per design point, reference run.py:67)."""
        with open(self.output_path_fnx, "ab+") as f:
            for rec in records:
                pickle.dump(rec, f)
