"""Character model store on the GPU.

Mirror of the reference's ``ModelManager`` / ``GSVModel`` / ``GSVModelFile``
(src/genie_tts/ModelManager.py:26-324): same file names, same LRU behaviour
(env ``Max_Cached_Character_Models``, default 3), same ``get`` / lazy reload
semantics — but a character is ONE ``B200Model`` (weights resident in HBM once,
not five onnxruntime sessions with the T2S weights duplicated in fp32), and the
five ``GSVModel`` fields are stage handles onto it so the call sites in
``GENIE.tts`` keep their argument lists."""
from __future__ import annotations

import gc
import logging
import os
from dataclasses import dataclass
from typing import Dict, Optional

from .Utils.Utils import LRUCacheDict

logger = logging.getLogger(__name__)


class GSVModelFile:
    T2S_ENCODER_FP32 = "t2s_encoder_fp32.onnx"
    T2S_FIRST_STAGE_DECODER_FP32 = "t2s_first_stage_decoder_fp32.onnx"
    T2S_STAGE_DECODER_FP32 = "t2s_stage_decoder_fp32.onnx"
    T2S_DECODER_WEIGHT_FP16 = "t2s_shared_fp16.bin"
    VITS_FP32 = "vits_fp32.onnx"
    VITS_WEIGHT_FP16 = "vits_fp16.bin"
    PROMPT_ENCODER = "prompt_encoder_fp32.onnx"
    PROMPT_ENCODER_WEIGHT_FP16 = "prompt_encoder_fp16.bin"


class StageSession:
    """Stands where an ``onnxruntime.InferenceSession`` stood in ``GSVModel``.
    ``run`` is deliberately absent: the per-step session.run contract (48 KV arrays through
    numpy, Inference.py:98-103) is what this build deletes; the host calls the engine's
    stage-group entry points on ``.model`` instead."""

    def __init__(self, model, role: str):
        self.model = model
        self.role = role

    def __repr__(self) -> str:
        return f"<StageSession {self.role} on cuda:{self.model.device}>"


@dataclass
class GSVModel:
    LANGUAGE: str
    T2S_ENCODER: StageSession
    T2S_FIRST_STAGE_DECODER: StageSession
    T2S_STAGE_DECODER: StageSession
    VITS: StageSession
    PROMPT_ENCODER: Optional[StageSession] = None
    PROMPT_ENCODER_PATH: Optional[str] = None

    @property
    def engine(self):
        return self.VITS.model


class ModelManager:
    def __init__(self, device: Optional[int] = None):
        cap = int(os.getenv("Max_Cached_Character_Models", "3"))
        # an evicted character releases its HBM at once: the model closes its device prompts and itself (a cached
        # ReferenceAudio rebuilds its device prompt on the next use, ``get`` reloads the weights lazily)
        self.character_to_model: Dict[str, object] = LRUCacheDict(capacity=cap, on_evict=self._evicted)
        self.character_to_language: Dict[str, str] = {}
        self.character_model_paths: Dict[str, str] = {}
        self.device = int(os.getenv("GENIE_DEVICE", "0")) if device is None else device
        self.providers = ["B200ExecutionProvider"]     # the reference pins CPUExecutionProvider (:125)

    @staticmethod
    def _evicted(name: str, model) -> None:
        logger.info(f"Character {name.capitalize()} evicted from the model cache.")
        try:
            model.close()
        except Exception as e:                     # never let an eviction break the load that caused it
            logger.error(f"Error while closing evicted character {name}: {e}")

    def load_character(self, character_name: str, model_dir: str, language: str) -> bool:
        name = character_name.lower()
        if name in self.character_to_model:
            _ = self.character_to_model[name]      # touch
            return True
        try:
            from .engine import B200Model
            model = B200Model(model_dir, device=self.device)
        except Exception as e:                     # same contract as the reference: log + False
            logger.error(f"Error: Failed to load model directory '{model_dir}'.\nDetails: {e}")
            return False
        self.character_to_model[name] = model      # may evict (and close) the least recently used character
        self.character_to_language[name] = language
        self.character_model_paths[name] = model_dir
        logger.info(f"Character {name.capitalize()} loaded successfully.\n- Model Path: {model_dir}\n"
                    f"- Model Type: {'V2ProPlus' if model.is_v2pp else 'V2'}")
        return True

    def get(self, character_name: str) -> Optional[GSVModel]:
        name = character_name.lower()
        language = self.character_to_language.get(name, "Japanese")
        if name in self.character_to_model:
            model = self.character_to_model[name]
            pe = StageSession(model, "prompt_encoder") if model.is_v2pp else None
            return GSVModel(
                LANGUAGE=language,
                T2S_ENCODER=StageSession(model, "t2s_encoder"),
                T2S_FIRST_STAGE_DECODER=StageSession(model, "t2s_first_stage_decoder"),
                T2S_STAGE_DECODER=StageSession(model, "t2s_stage_decoder"),
                VITS=StageSession(model, "vits"),
                PROMPT_ENCODER=pe,
                PROMPT_ENCODER_PATH=os.path.join(self.character_model_paths[name], GSVModelFile.PROMPT_ENCODER),
            )
        if name in self.character_model_paths:     # evicted: reload lazily (reference :218-224)
            if self.load_character(name, self.character_model_paths[name], language=language):
                return self.get(name)
            del self.character_model_paths[name]
        return None

    def has_character(self, character_name: str) -> bool:
        return character_name.lower() in self.character_model_paths

    def remove_character(self, character_name: str) -> None:
        name = character_name.lower()
        if name in self.character_to_model:
            model = self.character_to_model[name]
            del self.character_to_model[name]
            model.close()                          # waits for a call in flight, closes the model's device prompts
            gc.collect()
            logger.info(f"Character {name.capitalize()} removed successfully.")

    def remove_all_character(self) -> None:
        for m in list(self.character_to_model.values()):
            m.close()
        self.character_to_model.clear()
        gc.collect()


model_manager: ModelManager = ModelManager()
