"""Host-side mirrors of the reference's interfaces for the hot path (file:line citations live in each module's docstring).

    registry        string-keyed op registries + install_into_lightx2v()           <- lightx2v/utils/registry_factory.py
    ops             MMWeightB200 / MMWeightFp8B200 / MMWeightNvfp4B200, RMS / LN / FMHA / tensor ops, offline weight quantisers
    weight_module   WeightModule / WeightModuleList                                <- lightx2v/common/modules/weight_module.py
    wan_weights     WanTransformerWeights tree (checkpoint key names)              <- wan/weights/transformer_weights.py
    wan_infer       WanTransformerInfer (fused block schedule, native block driver) <- wan/infer/transformer_infer.py
    wan_teacache    WanTransformerInferTeaCaching                                  <- wan/infer/feature_caching/transformer_infer.py
    wan_causvid     WanTransformerInferCausVid (KV-cache block variant)            <- wan/infer/causvid/transformer_infer.py
    wan_model       WanPreInfer / WanPostInfer / WanModel (+ CFG-parallel)         <- wan/infer/pre_infer.py, post_infer.py, wan/model.py
    wan_scheduler   WanScheduler (UniPC), WanStepDistillScheduler                  <- schedulers/wan/scheduler.py, step_distill/scheduler.py
    ulysses         Ulysses SP: NCCL and peer-memory paths, Hunyuan variant, CFG-parallel <- attentions/distributed/ulysses/*
    wan_vae         WanVAEDecoderB200 (.decode / .decode_dist)                     <- video_encoders/hf/wan/vae.py
    hunyuan_infer   HunyuanTransformerWeights / HunyuanTransformerInfer            <- hunyuan/infer/transformer_infer.py
    hunyuan_vae     HunyuanVAEB200 (.decode / .decode_dist)                        <- video_encoders/hf/autoencoder_kl_causal_3d/*
"""
