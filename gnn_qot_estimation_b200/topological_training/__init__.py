from .models import TopologicalGNN  # noqa: F401
