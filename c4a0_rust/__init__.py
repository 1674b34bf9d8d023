"""Drop-in for the reference's `c4a0_rust` PyO3 module (rust/src/lib.rs:23-41,
src/c4a0_rust/__init__.pyi): same constants, classes and functions, backed by the B200 engine.

    from c4a0_rust import play_games, GameMetadata, PlayGamesResult, N_COLS, N_ROWS

`run_tui` (terminal UI) and `PlayGamesResult.score_policies` (external solver binary) are outside
the self-play hot path and raise NotImplementedError.
"""

from ._native import (  # noqa: F401
    BUF_N_CHANNELS,
    N_COLS,
    N_ROWS,
    GameMetadata,
    GameResult,
    PlayGamesResult,
    Sample,
    play_games,
    run_tui,
)

__all__ = [
    "BUF_N_CHANNELS", "N_COLS", "N_ROWS", "GameMetadata", "GameResult", "PlayGamesResult", "Sample", "play_games",
    "run_tui",
]
