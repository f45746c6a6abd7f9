"""Two-symbol stand-in for diffusers==0.24.0 so /root/reference/adapter/attention_processor.py imports unmodified
(its only diffusers imports are diffusers.utils.USE_PEFT_BACKEND and diffusers.models.lora.LoRALinearLayer,
attention_processor.py:6-7). Used ONLY by oracle/make_golden.py in the build container."""
