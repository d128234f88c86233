"""Host-side file discovery of the reference's ``utils/video/mov_extraction.py`` (no arithmetic).

``find_files`` keeps the reference's return tuple; ``get_audio`` returns the take's WAV, extracting
it from a ``.mov`` / ``.mp4`` with ffmpeg (``-ac 1 -ar sr``) when one is present and ffmpeg exists.
Out of the hot path's scope - kept only so the dataset builders run.
"""
import os
import shutil
import subprocess

DEFAULT_SR = 88200


def find_files(folder_path):
    """reference :8-29 -> (mov, mp4, wav, facial_csv, audio_features_csv, other_csv)."""
    found = {"mov": None, "mp4": None, "wav": None, "facial": None, "other": None}
    for name in os.listdir(folder_path):
        path = os.path.join(folder_path, name)
        ext = os.path.splitext(name)[1]
        if ext in (".mov", ".mp4", ".wav"):
            found[ext[1:]] = path
        elif ext == ".csv":
            found["facial" if "iPhone_cal" in name else "other"] = path
    cache = os.path.join(folder_path, "audio_features.csv")   # returned whether or not it exists
    return found["mov"], found["mp4"], found["wav"], found["facial"], cache, found["other"]


def extract_audio(video_path, output_dir, sr=DEFAULT_SR, ffmpeg_path=None):
    """reference :39-63 -- ffmpeg -i video -ac 1 -ar sr -y audio.wav; None on failure."""
    audio_path = os.path.join(output_dir, "audio.wav")
    if os.path.exists(audio_path):
        print(f"Audio already exists at {audio_path}")
        return audio_path
    ffmpeg = ffmpeg_path or os.environ.get("FFMPEG_PATH") or shutil.which("ffmpeg")
    if not ffmpeg:
        print(f"Failed to extract audio from {video_path}: ffmpeg not found")
        return None
    try:
        subprocess.run([ffmpeg, "-i", video_path, "-ac", "1", "-ar", str(sr), "-y", audio_path],
                       check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        return audio_path
    except subprocess.CalledProcessError as e:
        print(f"Failed to extract audio from {video_path}: {e.stderr.decode('utf-8', 'replace')}")
        return None


def get_audio(video_path, wav_path, folder_path):
    """reference :31-37."""
    return extract_audio(video_path, folder_path) if video_path else wav_path
