from p1afempy_b200.cubature import get_rule  # noqa: F401
