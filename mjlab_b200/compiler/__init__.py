from mjlab_b200.compiler.compile import Model, compile_spec
from mjlab_b200.compiler.spec import Spec

__all__ = ["Model", "Spec", "compile_spec"]
