"""Host-side mirrors of small helper tables defined in the CUDA sources (kept in sync by
tests/test_host_logic.py::test_stage_input_table_matches_cuda_source)."""


def etdrk_stage_input(order: int, s: int) -> int:
    """Which buffer feeds the nonlinear function of ETDRK stage s (-1: the step input, otherwise an
    index into the scratch states) -- exb_nl.cuh: etdrk_stage_input."""
    if s == 0:
        return -1
    if order == 4 and s >= 2:
        return 2
    return 0
