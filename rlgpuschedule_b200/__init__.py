"""B200-native hot path of the RLGPUSchedule cluster simulator."""
from .ingest import Cluster, Trace, cluster_from_flags, prepare_trace  # noqa: F401
from .simulator import Simulator, replay_event_rows  # noqa: F401
