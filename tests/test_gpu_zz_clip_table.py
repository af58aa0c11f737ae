"""GPU: slr_clip_table / slr_clip_bin / slr_clip_stats_host through the C ABI on device buffers.
The body (tests/clip_abi_cases.py) is shared with the CPU emulation suite."""
import pytest

pytestmark = pytest.mark.gpu


def test_clip_table_bin_stats_through_the_c_abi():
    import __graft_entry__
    __graft_entry__.build()
    import clip_abi_cases
    clip_abi_cases.clip_table_bin_stats(clip_abi_cases.CudaBackend())
