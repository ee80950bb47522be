from .arguments import TrainingArgs, parse_args  # noqa: F401
from .trainer import Trainer, lr_at_step, split_batch_by_props  # noqa: F401
