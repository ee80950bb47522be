from .synthetic import SyntheticCaptionTask, SyntheticControlTask, SyntheticTextTask, SyntheticVqaTask, build_synthetic_tasks  # noqa: F401
