"""REST wrapper over `Engine` with the reference's wire format (server/main.py:14-112 of the reference):
GET /legal_moves/{idx}, POST /play_move, POST /play_mcts, POST /add_game, GET /state/{idx}.
`/play_mcts` runs the batched CUDA search; everything else is host bookkeeping."""
import argparse
import collections
import os
from contextlib import asynccontextmanager
from typing import Any, Optional

from fastapi import FastAPI, HTTPException
from pydantic import BaseModel

from engine.engine import Engine


def serialize_state(state: object) -> dict:
    """board as a list; every public int/float/str/bool attribute by name; deques as lists
    (server/main.py:14-36).  Works for the chess `State` (nine fields, bindings_chess.cpp:13-41) and the
    Connect Four namedtuple alike."""
    out: dict = {}
    if hasattr(state, "board"):
        board = state.board
        try:
            out["board"] = list(board)
        except TypeError:
            out["board"] = board
    for name in dir(state):
        if name.startswith("_") or name == "board":
            continue
        try:
            value = getattr(state, name)
        except Exception:
            continue
        if isinstance(value, (bool, int, float, str)):
            out[name] = value
        elif isinstance(value, collections.deque):
            out[name] = list(value)
    return out


config_path: Optional[str] = None


@asynccontextmanager
async def lifespan(app: FastAPI):
    cfg = config_path or os.getenv("CONFIG_PATH")
    if not cfg:
        raise RuntimeError("Config file path must be set: `python -m server.main -c <yaml>` or CONFIG_PATH")
    app.state.engine = Engine(cfg)
    yield


app = FastAPI(lifespan=lifespan)


class MoveRequest(BaseModel):
    idx: int = 0
    move: Any


class MCTSRequest(BaseModel):
    idx: int = 0
    simulations: int = 1000
    c: float = 1.4


def _guard(fn):
    try:
        return fn()
    except HTTPException:
        raise
    except Exception as exc:      # the reference answers every failure with 400 + the message
        raise HTTPException(status_code=400, detail=str(exc))


@app.get("/legal_moves/{idx}")
def legal_moves(idx: int):
    return _guard(lambda: {"idx": idx, "moves": [list(m[0]) if isinstance(m[0], (tuple, list)) else [m[0]]
                                                 for m in app.state.engine.legal_moves(idx)]})


@app.post("/play_move")
def play_move(req: MoveRequest):
    def run():
        eng = app.state.engine
        wanted = tuple(req.move) if isinstance(req.move, (list, tuple)) else req.move
        legal = next((mv for mv in eng.legal_moves(req.idx)
                      if mv[0] == wanted or (isinstance(wanted, tuple) and len(wanted) == 1 and mv[0] == wanted[0])), None)
        if legal is None:
            raise ValueError("Illegal move")
        result = eng.play_move(legal, req.idx)
        return {"idx": req.idx, "result": result, **serialize_state(eng.get_state(req.idx))}
    return _guard(run)


@app.post("/play_mcts")
def play_mcts(req: MCTSRequest):
    def run():
        eng = app.state.engine
        result = eng.play_mcts(req.idx, req.simulations, req.c)
        return {"idx": req.idx, "result": result, **serialize_state(eng.get_state(req.idx))}
    return _guard(run)


@app.post("/add_game")
def add_game():
    return _guard(lambda: {"idx": app.state.engine.add_game()})


@app.get("/state/{idx}")
def get_state(idx: int):
    return _guard(lambda: {"idx": idx, **serialize_state(app.state.engine.get_state(idx))})


def main() -> None:
    import uvicorn

    ap = argparse.ArgumentParser(description="REST server for the game engine")
    ap.add_argument("-c", "--config", required=True, help="game configuration YAML")
    ap.add_argument("--host", default="0.0.0.0")
    ap.add_argument("--port", type=int, default=8000)
    args = ap.parse_args()
    global config_path
    config_path = args.config
    os.environ["CONFIG_PATH"] = args.config
    uvicorn.run(app, host=args.host, port=args.port)


if __name__ == "__main__":
    main()
