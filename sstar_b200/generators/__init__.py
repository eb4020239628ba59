from .generic_generator import GenericGenerator  # noqa: F401
from .random_number_generator import RandomNumberGenerator  # noqa: F401
from .window_data_generator import WindowDataGenerator  # noqa: F401
