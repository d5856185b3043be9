"""Import stub (test infrastructure): PESQ is a CPU metric outside the hot path (network_functions.py:7)."""


def pesq(*a, **k):
    return float("nan")
