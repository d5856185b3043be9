"""Import stub (test infrastructure): STOI is a CPU metric outside the hot path (network_functions.py:11)."""


def stoi(*a, **k):
    return float("nan")
