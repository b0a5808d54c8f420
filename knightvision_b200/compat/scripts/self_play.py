"""scripts.self_play of the reference tree -> the B200 self-play engine."""
from knightvision_b200.selfplay import generate_self_play_data, self_play  # noqa: F401
