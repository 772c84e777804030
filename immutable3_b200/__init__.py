"""immutable3_b200 — B200-native scan / filter / project path of immutable3 behind the reference's
Table / SegmentManager / Query / Engine interface.  Importing this package loads the CUDA shared
library lazily on first use and fails loudly if it is missing (there is no CPU fallback)."""
from .engine import (And, Avg, Column, Count, EQ, Engine, GT, LT, Match, Max, Min, NoOp, NoSelect, NotMatch, Or, Project, ProjectAgg, Query,
                     Result, Row, SegmentManager, Select, Sum, Table, flatten_select, select_dnf)
from ._lib import (CODEC_DENSE_INT, CODEC_DENSE_STRING, CODEC_DENSE_TINYINT, CODEC_PFOR_INT, Imm3Error, OPEN_FORCE_BLOCKS,
                   OPEN_HOST_ONLY, OPEN_KEEP_HOST, OPEN_NO_STATS, OPEN_NO_TMA)

__all__ = [n for n in dir() if not n.startswith("_")]
