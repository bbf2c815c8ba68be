from .FALoss import FALoss  # noqa: F401
from .CrossEntropyLoss import CrossEntropyLoss  # noqa: F401
