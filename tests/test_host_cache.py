"""Host logic of the drop-in blocks: the static-plan cache is keyed on LIVE index tensors (a freed index tensor's
address can be handed to a new tensor with different contents by the caching allocator)."""
import gc

import torch


def test_plan_cache_hits_while_key_tensor_is_alive():
    from graphs4cfd_b200.blocks import _Cache
    c = _Cache()
    idx = torch.arange(12)
    builds = []
    first = c.get(("mp", 4) + _Cache.key(idx), lambda: builds.append(1) or "plan-A")
    again = c.get(("mp", 4) + _Cache.key(idx), lambda: builds.append(1) or "plan-B")
    assert first == again == "plan-A" and len(builds) == 1
    other = c.get(("mp", 5) + _Cache.key(idx), lambda: "plan-C")          # different static part of the key
    assert other == "plan-C"


def test_plan_cache_rebuilds_when_key_tensor_died():
    from graphs4cfd_b200.blocks import _Cache
    c = _Cache()
    idx = torch.arange(12)
    ident = _Cache.key(idx).ident
    assert c.get(("mp", 4) + _Cache.key(idx), lambda: "old") == "old"
    del idx
    gc.collect()
    new = torch.arange(12) + 1                     # same shape; pretend the allocator reused the address
    key = _Cache.key(new)
    key.ident = ident
    assert c.get(("mp", 4) + key, lambda: "new") == "new"


def test_plan_cache_two_tensor_key():
    from graphs4cfd_b200.blocks import _Cache
    c = _Cache()
    a, b = torch.arange(3), torch.arange(5)
    assert c.get(("pool",) + _Cache.key(a, b), lambda: 1) == 1
    assert c.get(("pool",) + _Cache.key(a, b), lambda: 2) == 1
    assert c.get(("pool",) + _Cache.key(b, a), lambda: 3) == 3
