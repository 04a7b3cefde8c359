"""Parallelism strategies: the data-parallel engine, its algorithms, expert parallelism (MoE) and the NVSwitch
symmetric-memory engine."""
