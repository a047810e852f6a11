// nlc_model_forward_ts lives in rollout.cu (it shares the fused MLP/ILT kernel).
