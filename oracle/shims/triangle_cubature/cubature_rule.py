from p1afempy_b200.cubature import CubatureRuleEnum  # noqa: F401
