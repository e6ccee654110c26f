"""`-m gpu`: the whole CUDA forward (XceptionVidTr.forward -> C ABI) against the golden vectors produced from the
unmodified reference, and against the CPU oracle."""
import pytest

import model_checks


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("case", ["default_init_b1", "sensitised_b2", "sensitised_t32_b1"])
def test_golden(case, precision):
    model_checks.run_golden_case(case, precision)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_against_cpu_oracle(precision):
    model_checks.run_oracle_case(precision)


@pytest.mark.gpu
def test_benchmark_shape_batch64_against_oracle():
    model_checks.run_batch64_parity()


@pytest.mark.gpu
def test_data_parallel_two_devices():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("nn.DataParallel check needs two visible GPUs (run under gpurun --gpus 2)")
    model_checks.run_data_parallel_check()


@pytest.mark.gpu
def test_api_boundary():
    model_checks.run_api_checks()


@pytest.mark.gpu
def test_training_step_against_reference_golden():
    model_checks.run_train_golden()


@pytest.mark.gpu
def test_training_step_long_clip_against_oracle():
    model_checks.run_train_t32_oracle()


@pytest.mark.gpu
def test_relevance_pass_against_oracle():
    model_checks.run_relevance_check()


@pytest.mark.gpu
def test_relevance_pass_long_clip_against_oracle():
    model_checks.run_relevance_t32_check()


@pytest.mark.gpu
def test_cuda_graph_forward():
    model_checks.run_graph_check()


@pytest.mark.gpu
def test_persisted_pack_cache():
    model_checks.run_pack_cache_check()


@pytest.mark.gpu
def test_uint8_frames_against_oracle():
    model_checks.run_uint8_input_check()


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_xception_baseline_against_reference_golden(precision):
    model_checks.run_xception_golden(precision)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("case", ["vivit_d2_b2", "vivit_mean_d2_b2", "vivit_d12_b1", "vanilla_d2_b1", "vanilla_d12_b1"])
def test_ablation_models_against_reference_golden(case, precision):
    model_checks.run_ablation_golden(case, precision)


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_ablation_blocks_against_reference_golden(precision):
    model_checks.run_ablation_blocks(precision)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["vivit", "vanilla"])
def test_ablation_variants_behind_the_entry_flow(variant):
    model_checks.run_ablation_clip_check(variant)
