"""FIR coefficient sets of the reference (feng/ddc/src/ddc_coeff_107MHz.csv, ddc_coeff_53MHz.csv).

The reference ships 256 symmetric low-pass taps per file; every value is k / 2**17 printed with '%.5g'.  This
package stores the integer numerators (data/taps_q17.json) and regenerates the text files on demand so that
`DigitalDownConverter(..., ddc_coeff_filename="…/ddc_coeff_107MHz.csv")` works exactly as with the reference.
"""
from __future__ import annotations

import json
import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "taps_q17.json")
NAMES = ("ddc_coeff_107MHz.csv", "ddc_coeff_53MHz.csv")


def numerators(name: str) -> list[int]:
    with open(_DATA) as f:
        return json.load(f)[name]


def csv_text(name: str) -> str:
    return "".join("%.5g\n" % (k / 131072.0) for k in numerators(name))


def coefficients(name: str) -> np.ndarray:
    """float64 taps exactly as numpy.genfromtxt parses the reference file (ddc.py:46)."""
    return np.array([float("%.5g" % (k / 131072.0)) for k in numerators(name)], dtype=np.float64)


def write_csv(name: str, directory: str) -> str:
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, name)
    with open(path, "w") as f:
        f.write(csv_text(name))
    return path


def write_all(directory: str) -> list[str]:
    return [write_csv(n, directory) for n in NAMES]
