BERT_FEATURE_DIM = 1024      # width of text_bert rows (reference Utils/Constants.py:1)
PACKAGE_NAME = "genie_tts"
SAMPLE_RATE = 32000          # output sampling rate (reference Core/TTSPlayer.py:25)
SAMPLES_PER_TOKEN = 1280     # 25 Hz semantic tokens -> 32 kHz audio
