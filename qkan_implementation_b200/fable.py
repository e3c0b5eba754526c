"""Placeholder until the FABLE gate-list generator lands."""


def fable(matrix, eps=0):
    raise NotImplementedError("FABLE circuit construction is not built yet (next row of the scope table)")
