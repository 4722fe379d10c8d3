"""CPU: bookkeeping of _capi.HostPins (page-locked registrations of caller-owned feed arrays) against a stand-in for the two
C-ABI calls; the real registration is exercised on the GPU (tests/test_gpu_parity.py)."""
import numpy as np

from varnet_b200 import _capi


class FakeLib:
    def __init__(self, rc=0):
        self.rc, self.reg, self.unreg = rc, [], []

    def vn_host_register(self, p, n):
        self.reg.append((p.value, n))
        return self.rc

    def vn_host_unregister(self, p):
        self.unreg.append(p.value)
        return 0


def make(monkeypatch, rc=0, **kw):
    lib = FakeLib(rc)
    monkeypatch.setattr(_capi, "load_library", lambda path=None: lib)
    return lib, _capi.HostPins(**kw)


def test_registers_on_the_second_sighting_and_releases(monkeypatch):
    lib, pins = make(monkeypatch, cap_bytes=1 << 20, min_bytes=1000)
    a, b = np.zeros((100, 3)), np.zeros((100, 1))
    assert not pins.touch_group([a, None, b])                       # first sighting
    assert lib.reg == []
    assert pins.touch_group([a, None, b]) and pins.registered == 2 and pins.bytes == a.nbytes + b.nbytes
    assert pins.touch_group([a, None, b]) and len(lib.reg) == 2     # cached
    pins.release()
    assert sorted(lib.unreg) == sorted(k for k, _ in lib.reg) and pins.bytes == 0


def test_small_mixed_or_non_contiguous_groups_are_left_alone(monkeypatch):
    lib, pins = make(monkeypatch, cap_bytes=1 << 20, min_bytes=1000)
    small = np.zeros(10)
    big64, big32 = np.zeros(1000), np.zeros(1000, dtype=np.float32)
    for _ in range(3):
        assert not pins.touch_group([small])
        assert not pins.touch_group([big64, big32])                 # one dtype per group (the engine call casts the others: copies)
        assert not pins.touch_group([np.zeros((100, 20))[:, ::2]])
        assert not pins.touch_group([[1.0, 2.0]])
        assert not pins.touch_group([np.array([[None]], dtype=object)])
    assert lib.reg == []


def test_eviction_is_lru_and_never_thrashes(monkeypatch):
    lib, pins = make(monkeypatch, cap_bytes=3000 * 8, min_bytes=1000)
    synced = []
    pins.sync = lambda: synced.append(1)
    arrs = [np.zeros(1000) for _ in range(4)]
    for a in arrs[:3]:
        pins.touch_group([a]); assert pins.touch_group([a])
    assert pins.bytes == 3000 * 8
    # a fourth array while the first three were in use a moment ago: refused, nothing evicted
    pins.touch_group([arrs[3]])
    assert not pins.touch_group([arrs[3]]) and lib.unreg == []
    # the first one falls out of use: the fourth may take its place (after a device sync)
    for _ in range(12):
        assert pins.touch_group([arrs[1]]) and pins.touch_group([arrs[2]])
    assert pins.touch_group([arrs[3]])
    assert lib.unreg == [arrs[0].ctypes.data] and synced and pins.bytes == 3000 * 8


def test_memory_locked_by_its_owner_is_used_but_never_unregistered(monkeypatch):
    lib, pins = make(monkeypatch, rc=1, cap_bytes=1 << 20, min_bytes=1000)
    a = np.zeros(1000)
    pins.touch_group([a])
    assert pins.touch_group([a]) and pins.registered == 0 and pins.bytes == 0
    pins.release()
    assert lib.unreg == []


def test_a_refused_registration_switches_the_registry_off(monkeypatch):
    lib, pins = make(monkeypatch, rc=3, cap_bytes=1 << 20, min_bytes=1000)
    a = np.zeros(1000)
    pins.touch_group([a])
    assert not pins.touch_group([a]) and pins.failed
    assert not pins.touch_group([a]) and len(lib.reg) == 1


def test_a_recycled_address_is_not_a_second_sighting(monkeypatch):
    """A feed that is rebuilt for every step (new arrays, the old ones freed) often lands on the same address: never registered."""
    lib, pins = make(monkeypatch, cap_bytes=1 << 30, min_bytes=1000)
    base = np.zeros(4000)

    class View(np.ndarray):                           # distinct objects over the same memory = "a new array at the old address"
        pass

    for _ in range(4):
        assert not pins.touch_group([base[:2000].view(View)])
    assert lib.reg == []
    keep = base[:2000].view(View)
    assert not pins.touch_group([keep])
    assert pins.touch_group([keep]) and len(lib.reg) == 1
