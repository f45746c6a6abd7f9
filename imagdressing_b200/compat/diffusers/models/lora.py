from adapter.attention_processor import LoRALinearLayer  # noqa: F401
