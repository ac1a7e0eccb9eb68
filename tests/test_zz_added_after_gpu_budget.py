"""GPU twins of CPU-pinned CLI goldens (configs[3] with its own sample count, configs[1] with the annotation).  Added after
the GPU budget of round 1 was spent, so this module collects last: its first run on a B200 cannot stop the earlier parity
modules under `pytest -x`."""
import pytest

pytestmark = pytest.mark.gpu


def test_c4_with_48_samples_equals_the_reference_files(ctx, tmp_path):
    """configs[3]'s own sample count (48 = 6 conditions x 8 replicates, 1.44M records, 24k re-counted gaps in `combine`): the
    CLI with the CUDA path must write the bytes the unmodified reference wrote (oracle/c4_shape.py c4x48)."""
    import json
    from oracle import c4_shape
    from spliser_b200 import cli
    shape = c4_shape.FORTY_EIGHT
    gold = json.load(open(shape.golden))
    per_sample, combined, shallow = c4_shape.run_cli(cli, ctx, str(tmp_path), shape)
    assert per_sample == gold["process_sha256"]
    assert combined == gold["combined_sha256"]
    assert shallow == gold["shallow_sha256"]                  # combineShallow -m 24 -r 4 -e 0.05 over the same samples


def test_annotated_process_equals_the_reference_files(ctx, tmp_path):
    """configs[1]'s flow with the GFF annotation at reduced size (stranded, unstranded, shuffled annotation, -g / -c / -m): the CLI
    with the CUDA path must write the bytes the unmodified reference wrote, Gene column included (oracle/c2_annotated.py)."""
    import json
    from oracle import c2_annotated
    from spliser_b200 import cli
    gold = json.load(open(c2_annotated.GOLDEN))["variants"]
    got = c2_annotated.run_cli(cli, ctx, str(tmp_path))
    assert sorted(got) == sorted(gold)
    for name, digest in got.items():
        assert digest == gold[name]["sha256"], name
