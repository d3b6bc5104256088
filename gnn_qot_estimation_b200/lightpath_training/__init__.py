from .models import LightpathGNN  # noqa: F401
