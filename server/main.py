"""HTTP front of `Engine` speaking the reference's wire format (its server/main.py:14-112): five routes --
legal moves, play a move, let the search play, open a game, read a state -- and the state serialiser the PHP
board expects.  Only `/play_mcts` touches the GPU (one batched CUDA search); the rest is host bookkeeping.

    python -m server.main -c configs/crude_chess.yaml --port 8000
    CONFIG_PATH=configs/connect4.yaml uvicorn server.main:app
"""
import argparse
import os
from collections import deque
from contextlib import asynccontextmanager
from typing import Any, Callable, Optional

from fastapi import APIRouter, FastAPI, HTTPException, Request
from pydantic import BaseModel, Field

from engine.engine import Engine

_SCALARS = (bool, int, float, str)


def serialize_state(state: object) -> dict:
    """Wire form of a backend state: `board` as a list when it can be listed, then every public attribute that is
    a plain scalar, deques flattened to lists.  (pybind11 hands the chess histories over as lists, which the
    reference's serialiser skips -- so they are not part of the format, here as there.)"""
    wire: dict = {}
    board = getattr(state, "board", None)
    if hasattr(state, "board"):
        try:
            wire["board"] = list(board)
        except TypeError:
            wire["board"] = board
    for attr in (a for a in dir(state) if a != "board" and not a.startswith("_")):
        try:
            value = getattr(state, attr)
        except Exception:
            continue
        if isinstance(value, deque):
            wire[attr] = list(value)
        elif isinstance(value, _SCALARS):
            wire[attr] = value
    return wire


class MoveRequest(BaseModel):
    idx: int = 0
    move: Any


class MCTSRequest(BaseModel):
    idx: int = 0
    # bounded: the search handle's node arena is sized from this number (sims+1 nodes per tree)
    simulations: int = Field(1000, ge=1, le=200_000)
    c: float = 1.4


config_path: Optional[str] = None      # set by main(); CONFIG_PATH is the alternative
routes = APIRouter()


def _engine(request: Request) -> Engine:
    return request.app.state.engine


def _answer(work: Callable[[], dict]) -> dict:
    """every failure is a 400 carrying the exception text, as in the reference"""
    try:
        return work()
    except HTTPException:
        raise
    except Exception as exc:
        raise HTTPException(status_code=400, detail=str(exc))


def _with_state(eng: Engine, idx: int, **head) -> dict:
    return {"idx": idx, **head, **serialize_state(eng.get_state(idx))}


def _coords(move) -> list:
    first = move[0]
    return list(first) if isinstance(first, (tuple, list)) else [first]


@routes.get("/legal_moves/{idx}")
def legal_moves(idx: int, request: Request):
    return _answer(lambda: {"idx": idx, "moves": [_coords(m) for m in _engine(request).legal_moves(idx)]})


@routes.post("/play_move")
def play_move(req: MoveRequest, request: Request):
    def work():
        eng = _engine(request)
        wanted = req.move
        key = tuple(wanted) if isinstance(wanted, (list, tuple)) else (wanted,)
        for candidate in eng.legal_moves(req.idx):
            if tuple(_coords(candidate)) == key:
                return _with_state(eng, req.idx, result=eng.play_move(candidate, req.idx))
        raise ValueError("Illegal move")
    return _answer(work)


@routes.post("/play_mcts")
def play_mcts(req: MCTSRequest, request: Request):
    def work():
        eng = _engine(request)
        return _with_state(eng, req.idx, result=eng.play_mcts(req.idx, req.simulations, req.c))
    return _answer(work)


@routes.post("/add_game")
def add_game(request: Request):
    return _answer(lambda: {"idx": _engine(request).add_game()})


@routes.get("/state/{idx}")
def get_state(idx: int, request: Request):
    return _answer(lambda: _with_state(_engine(request), idx))


@asynccontextmanager
async def _lifespan(application: FastAPI):
    chosen = config_path or os.getenv("CONFIG_PATH")
    if not chosen:
        raise RuntimeError("Config file path must be set: `python -m server.main -c <yaml>` or CONFIG_PATH")
    application.state.engine = Engine(chosen)
    yield


app = FastAPI(lifespan=_lifespan)
app.include_router(routes)


def main() -> None:
    import uvicorn

    cli = argparse.ArgumentParser(description="REST server for the game engine")
    cli.add_argument("-c", "--config", required=True, help="game configuration YAML")
    cli.add_argument("--host", default="0.0.0.0")
    cli.add_argument("--port", type=int, default=8000)
    opts = cli.parse_args()
    global config_path
    config_path = opts.config
    os.environ["CONFIG_PATH"] = opts.config
    uvicorn.run(app, host=opts.host, port=opts.port)


if __name__ == "__main__":
    main()
