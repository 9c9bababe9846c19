"""Constants shared by the plugins. The prompt-related values are part of the model's training recipe, so they must equal the
reference's (t2v_metrics/constants.py:1-8) character for character; `tests/test_host_logic.py` compares them with the reference's
module when it is available."""

# where tokenizers / processors are cached when a plugin is allowed to download them (never the case in this sandbox)
HF_CACHE_DIR = "./hf_cache/"

# CLIP-FlanT5 (LLaVA-1.5-style) prompt assembly -------------------------------------------------------------------------------------
# longest token sequence the T5 tokenizer of the wrapper accepts (model_max_length of the v3.0 loader)
CONTEXT_LEN = 2048
# conversation header placed before " USER: <image>\n{question} ASSISTANT: " by format_question(..., 't5_chat')
SYSTEM_MSG = "A chat between a curious user and an artificial intelligence assistant. The assistant gives helpful, detailed, and polite answers to the user's questions."
# placeholder in the prompt text and the id it becomes in input_ids; the engine splices the 576 projected CLIP features there
DEFAULT_IMAGE_TOKEN = "<image>"
IMAGE_TOKEN_INDEX = -200
# label positions CrossEntropyLoss (and the engine's score reduction) skips
IGNORE_INDEX = -100
