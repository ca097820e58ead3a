from simwave_b200.kernel.backend.compiler import Compiler
from simwave_b200.kernel.backend.middleware import Middleware

__all__ = ["Compiler", "Middleware"]
