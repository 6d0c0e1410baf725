"""Synthetic-input producers for the BASELINE configs (shell tables, geometries, grids, density matrices).

These stand in for Serenity's BasisController / GridController / DensityMatrixController, which deliver the
inputs of the hot path; they are not part of it.
"""
from .basis import ShellTable, build_shell_table, shell_table_from_list  # noqa: F401
from .configs import make_config  # noqa: F401
