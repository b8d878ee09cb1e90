"""CPU: host logic of the shared search-handle pool (zeroclone_b200/mcts.py) with the device handle replaced by a stand-in --
capacities only grow, a failed growth leaves the pool consistent, one lock per pooled handle -- and of the `mcts:` config keys
that select the search mode."""
import threading

import pytest

from zeroclone_b200 import _ffi, mcts


class FakeSearch:
    created, closed, fail_above = [], [], None

    def __init__(self, game, max_trees, max_sims, device=0, arena_slots_per_tree=0):
        if FakeSearch.fail_above is not None and max_trees * max_sims > FakeSearch.fail_above:
            raise _ffi.ZcError(_ffi.ZC_ECUDA, "cudaMalloc: out of memory")
        self.game, self.max_trees, self.max_sims, self.device, self.open = game, max_trees, max_sims, device, True
        FakeSearch.created.append(self)

    def close(self):
        self.open = False
        FakeSearch.closed.append(self)


@pytest.fixture()
def pool(monkeypatch):
    monkeypatch.setattr(mcts, "TreeSearch", FakeSearch)
    monkeypatch.setattr(mcts, "_POOL", {})
    FakeSearch.created, FakeSearch.closed, FakeSearch.fail_above = [], [], None
    yield
    mcts._POOL.clear()


def test_capacities_only_grow(pool):
    a = mcts.searcher(_ffi.GAME_C4, 100, 800, device=0)
    assert (a.max_trees, a.max_sims) == (100, 800)
    assert mcts.searcher(_ffi.GAME_C4, 50, 400, device=0) is a                  # fits: same handle
    b = mcts.searcher(_ffi.GAME_C4, 10, 1600, device=0)                         # only sims grew: trees keep their capacity
    assert (b.max_trees, b.max_sims) == (100, 1600) and not a.open and b.open
    c = mcts.searcher(_ffi.GAME_C4, 150, 100, device=0)                         # trees grow geometrically, sims never shrink
    assert (c.max_trees, c.max_sims) == (200, 1600) and not b.open
    d = mcts.searcher(_ffi.GAME_CHESS, 7, 64, device=0)                         # another game: another handle
    assert d is not c and c.open and mcts.searcher(_ffi.GAME_C4, 1, 1, device=0) is c


def test_failed_growth_keeps_the_pool_usable(pool):
    a = mcts.searcher(_ffi.GAME_C4, 64, 100, device=0)
    FakeSearch.fail_above = 64 * 100                      # anything larger cannot be allocated
    with pytest.raises(_ffi.ZcError):
        mcts.searcher(_ffi.GAME_C4, 64, 10_000, device=0)
    # the old handle was given up to make room for the retry; the entry is gone rather than pointing at a closed handle
    assert not a.open and (_ffi.GAME_C4, 0) not in mcts._POOL
    b = mcts.searcher(_ffi.GAME_C4, 32, 100, device=0)
    assert b.open and (b.max_trees, b.max_sims) == (32, 100)
    # a growth that fails next to the old handle but fits once it is released succeeds on the retry
    FakeSearch.fail_above = 40 * 100
    c = mcts.searcher(_ffi.GAME_C4, 40, 100, device=0)
    assert c.open and not b.open and c.max_trees == 40 and mcts._POOL[(_ffi.GAME_C4, 0)] is c


def test_one_lock_per_pooled_handle(pool):
    la, lb = mcts.device_lock(_ffi.GAME_C4, 0), mcts.device_lock(_ffi.GAME_C4, 0)
    assert la is lb and mcts.device_lock(_ffi.GAME_CHESS, 0) is not la and mcts.device_lock(_ffi.GAME_C4, 1) is not la
    order = []

    def worker(tag):
        with mcts.device_lock(_ffi.GAME_C4, 0):
            order.append((tag, "in"))
            threading.Event().wait(0.05)
            order.append((tag, "out"))

    ts = [threading.Thread(target=worker, args=(i,)) for i in range(3)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert all(order[i][0] == order[i + 1][0] for i in range(0, 6, 2))          # critical sections never interleave


def test_select_mode_keys():
    assert mcts.select_mode(None) == (_ffi.SELECT_UCB1, 1.0, 0)
    assert mcts.select_mode({"simulations": 800}) == (_ffi.SELECT_UCB1, 1.0, 0)          # the reference's YAMLs: UCB1
    assert mcts.select_mode({"select": "PUCT", "virtual_loss": 0.5, "prior_weight": 2}) == (_ffi.SELECT_PUCT, 0.5, 2)
    with pytest.raises(ValueError):
        mcts.select_mode({"select": "minimax"})
