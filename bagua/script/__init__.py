"""Alias of :mod:`bagua_b200.script` (reference: bagua/script/: ``baguarun`` and ``bagua_sys_perf``)."""
