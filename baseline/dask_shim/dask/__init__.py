"""Stand-in for `dask` (the reference's only dependency missing from this image): the reference imports
`dask.delayed` / `dask.compute` at module level (quantum/fock_tensors.py:25, samples.py:57) but uses them only on its
`parallel=True` branches, which the baseline runs never take.  Put on sys.path by bench.py's reference arm only."""


def delayed(f, *a, **k):
    return f


def compute(*a, **k):
    return a
