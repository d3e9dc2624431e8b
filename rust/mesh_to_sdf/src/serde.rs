//! `mesh_to_sdf::serde` (feature `serde`): signed distance fields on disk, format V1 of the reference crate
//! (`mesh_to_sdf/src/serde.rs`): MessagePack via rmp-serde, enums as one-entry maps keyed by the variant name,
//! structs as arrays in field order —
//!   `{"V1": {"Generic": [[points...], [distances...]]}}`, `{"V1": {"Grid": [[first, size, count], [distances...]]}}`.
//! Host-only: nothing here touches the GPU. Files written by the reference crate load here and the other way round
//! (the Python and C++ mirrors in this repository are pinned byte for byte by the reference's fixtures).
use crate::{Grid, Point};
use serde::{de::DeserializeOwned, Deserialize, Serialize};
use std::path::Path;

#[derive(Debug)]
pub enum SerdeError {
    SerializationFailed(rmp_serde::encode::Error),
    DeserializationFailed(rmp_serde::decode::Error),
    IoError(std::io::Error),
}
impl From<std::io::Error> for SerdeError {
    fn from(e: std::io::Error) -> Self {
        Self::IoError(e)
    }
}
impl From<rmp_serde::encode::Error> for SerdeError {
    fn from(e: rmp_serde::encode::Error) -> Self {
        Self::SerializationFailed(e)
    }
}
impl From<rmp_serde::decode::Error> for SerdeError {
    fn from(e: rmp_serde::decode::Error) -> Self {
        Self::DeserializationFailed(e)
    }
}

#[derive(Serialize)]
#[serde(bound = "V: Serialize + DeserializeOwned")]
pub enum SerializeSdf<'a, V: Point> {
    Generic(SerializeGeneric<'a, V>),
    Grid(SerializeGrid<'a, V>),
}
#[derive(Serialize)]
#[serde(bound = "V: Serialize + DeserializeOwned")]
pub struct SerializeGeneric<'a, V: Point> {
    pub query_points: &'a [V],
    pub distances: &'a [f32],
}
#[derive(Serialize)]
#[serde(bound = "V: Serialize + DeserializeOwned")]
pub struct SerializeGrid<'a, V: Point> {
    pub grid: &'a Grid<V>,
    pub distances: &'a [f32],
}
#[derive(Serialize)]
#[serde(bound = "V: Serialize + DeserializeOwned")]
enum SerializeVersion<'a, V: Point> {
    V1(&'a SerializeSdf<'a, V>),
}

#[derive(Deserialize)]
#[serde(bound = "V: Serialize + DeserializeOwned")]
pub enum DeserializeSdf<V: Point> {
    Generic(DeserializeGeneric<V>),
    Grid(DeserializeGrid<V>),
}
#[derive(Deserialize)]
#[serde(bound = "V: Serialize + DeserializeOwned")]
pub struct DeserializeGeneric<V: Point> {
    pub query_points: Vec<V>,
    pub distances: Vec<f32>,
}
#[derive(Deserialize)]
#[serde(bound = "V: Serialize + DeserializeOwned")]
pub struct DeserializeGrid<V: Point> {
    pub grid: Grid<V>,
    pub distances: Vec<f32>,
}
#[derive(Deserialize)]
#[serde(bound = "V: Serialize + DeserializeOwned")]
enum DeserializeVersion<V: Point> {
    V1(DeserializeSdf<V>),
}

pub fn save_to_file<V: Point + Serialize + DeserializeOwned, P: AsRef<Path>>(sdf: &SerializeSdf<V>, path: P) -> Result<(), SerdeError> {
    std::fs::write(path, rmp_serde::to_vec(&SerializeVersion::V1(sdf))?)?;
    Ok(())
}

pub fn read_from_file<V: Point + Serialize + DeserializeOwned, P: AsRef<Path>>(path: P) -> Result<DeserializeSdf<V>, SerdeError> {
    let DeserializeVersion::V1(sdf) = rmp_serde::from_slice(&std::fs::read(path)?)?;
    Ok(sdf)
}
