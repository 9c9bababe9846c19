from .vqascore_models import list_all_vqascore_models, get_vqascore_model  # noqa: F401
