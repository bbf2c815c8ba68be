from .FALoss import FALoss  # noqa: F401
from .CrossEntropyLoss import CrossEntropyLoss  # noqa: F401
from .Stage3Loss import Stage3Loss  # noqa: F401
