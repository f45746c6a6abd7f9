USE_PEFT_BACKEND = False  # peft is not in the reference's requirements.txt
